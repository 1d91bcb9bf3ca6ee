"""Command-line feature extraction to .npy files: B200-path mirror of the reference's /extract_feature.py.

Same flags, same on-disk layouts (reference extract_feature.py:20-151, figures/output_format.jpg):

  default                 <output_dir>/<layer id>/<name>.npy      one (C, h, w) fp16 array per (layer, image)
  --sample_name_first     <output_dir>/<name>/<layer id>.npy
  --aggregate_output      <output_dir>/<name>.npy                 every map nearest-resized to the largest width and
                                                                  concatenated on channels (F.interpolate default mode,
                                                                  extract_feature.py:115-129)
  <name> = original file stem (--use_original_filename; with --nested_input_dir prefixed by the parent folder) or
  <split><running index>.

What is different underneath: images are decoded / resized by a prefetch thread while the GPU works on the previous
batch, the captured maps leave the GPU as ONE asynchronous device-to-host copy of the arena slots into pinned memory
on a side stream, and a pool of writer threads turns them into .npy files - disk and PIL never gate the kernels.
(The reference decodes, extracts, `.cpu()`s every map and np.save's them in one serial loop.)

    python -m generic_diffusion_feature_b200.extract_feature --layer feature/configs/config_xl_practical.json \
        --version xl --img_size 1024 --t 50 -b 8 --input_dir 'imgs/*.jpg' --prompt_file prompt.txt --output_dir out/
"""
import argparse
import concurrent.futures
import glob
import json
import os
import queue
import threading

import numpy as np
import torch


# flag table: (names, kwargs). Same flag names / defaults as the reference CLI (extract_feature.py:18-47) so that
# existing command lines keep working; descriptions are ours.
_FLAGS = [
    # model / capture selection
    (("--layer",), dict(type=str, help="JSON file (or nothing with --show_all_layers) naming the feature ids to keep")),
    (("--version",), dict(type=str, default="xl", help="xl | pgv2 | 2-1 | 1-5 | pixart-sigma | pixart-sigma-512 | flux")),
    (("--dtype",), dict(type=str, default="float16", choices=("float16", "float32"),
                        help="checked like the reference (models.py:11-16); the kernels compute in bf16 with fp32 "
                             "accumulation and the stored maps are fp16 for either value (FeatureStore casts them too)")),
    (("--offline_lora",), dict(type=str, default=None, help="not built on this path")),
    (("--offline_lora_filename",), dict(type=str, default=None, help="not built on this path")),
    (("--feature_resize",), dict(type=int, default=1, help="average-pool every stored map by this factor")),
    (("--control",), dict(type=str, nargs="+", default=None, help="not built on this path")),
    (("--attention",), dict(type=str, nargs="+", default=None,
                            choices=("down_cross", "mid_cross", "up_cross", "down_self", "mid_self", "up_self"),
                            help="attention categories aggregated into the `attn` feature")),
    (("--img_size",), dict(type=int, default=1024, help="images are resized to img_size x img_size")),
    # extraction
    (("--batch_size", "-b"), dict(type=int, default=2, help="images per extract call")),
    (("--t",), dict(type=int, help="q_sample timestep (0..1000)")),
    (("--denoising_from",), dict(type=int, default=None, help="not built on this path")),
    (("--use_ddim_inversion",), dict(action="store_true", help="not built on this path")),
    # input / output
    (("--input_dir",), dict(type=str, default=None, help="glob pattern of the input images")),
    (("--nested_input_dir",), dict(action="store_true", help="prefix output names with the image's parent folder")),
    (("--prompt_file",), dict(type=str, default="prompt.txt", help="text file holding the prompt")),
    (("--output_dir",), dict(type=str, default="./output/", help="root of the .npy tree")),
    (("--aggregate_output",), dict(action="store_true", help="one concatenated array per image")),
    (("--use_original_filename",), dict(action="store_true", help="name outputs after the input files")),
    (("--split",), dict(type=str, default="train", help="name prefix when not using original file names")),
    (("--sample_name_first",), dict(action="store_true", help="<out>/<name>/<layer>.npy instead of <out>/<layer>/<name>.npy")),
    (("--show_all_layers",), dict(action="store_true", help="print every available feature id and exit")),
    # B200-path extras
    (("--checkpoint",), dict(type=str, default=None,
                             help="diffusers-layout model directory (<dir>/unet|transformer + <dir>/vae, .safetensors); "
                                  "default: $GDF_MODEL_DIR[/<version>]")),
    (("--prompt_embeds",), dict(type=str, default=None,
                                help=".pt / .safetensors file holding the encoded prompt (the tuple encode_prompt returns, or "
                                     "a dict prompt_embeds / pooled_prompt_embeds / prompt_attention_mask); text encoders "
                                     "are not part of this path")),
    (("--synthetic",), dict(action="store_true",
                            help="explicit opt-in: random-init weights and/or stand-in prompt embeddings when no "
                                 "--checkpoint / --prompt_embeds is given (benchmarks, smoke tests)")),
    (("--writer_threads",), dict(type=int, default=8, help="threads writing .npy files")),
    (("--device",), dict(type=str, default="cuda", help="cuda | cuda:N (torchrun picks cuda:LOCAL_RANK)")),
]


def build_parser():
    parser = argparse.ArgumentParser(description="feature extraction to .npy files on the B200 path")
    for names, kw in _FLAGS:
        parser.add_argument(*names, **kw)
    return parser


def list_inputs(input_dir, nested):
    """[(path, save_name)] like the reference (:66-73)."""
    out = []
    for img in sorted(glob.glob(input_dir, recursive=True)):
        stem = os.path.splitext(os.path.basename(img))[0]
        name = stem if not nested else os.path.join(os.path.basename(os.path.split(img)[0]), stem)
        out.append((img, name))
    return out


def sample_name(args, dataset, index):
    return dataset[index][1] if args.use_original_filename else args.split + str(index)


def output_path(args, name, layer_id=None):
    """Path (without the .npy np.save appends) of one array, directories created like the reference (:131-147)."""
    if layer_id is None:                      # aggregated
        if args.nested_input_dir:
            os.makedirs(os.path.join(args.output_dir, name.split('/')[0]), exist_ok=True)
        else:
            os.makedirs(args.output_dir, exist_ok=True)
        return os.path.join(args.output_dir, name)
    if not args.sample_name_first:
        d = os.path.join(args.output_dir, layer_id)
        p = os.path.join(d, name)
    else:
        d = os.path.join(args.output_dir, name)
        p = os.path.join(d, layer_id)
    os.makedirs(os.path.dirname(p) if args.nested_input_dir else d, exist_ok=True)
    return p


def nearest_resize_chw(a, size):
    """F.interpolate(v, size) with its default mode='nearest' on one (C, h, w) array: src = floor(dst * in / out)."""
    c, h, w = a.shape
    if h == size and w == size:
        return a
    iy = np.minimum((np.arange(size) * (h / size)).astype(np.int64), h - 1)
    ix = np.minimum((np.arange(size) * (w / size)).astype(np.int64), w - 1)
    return a[:, iy][:, :, ix]


class NpyWriter:
    """Writer pool: takes host (pinned) token-major maps of one batch and writes the reference's .npy layout."""

    def __init__(self, args, dataset, threads):
        self.args, self.dataset = args, dataset
        self.pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, threads))
        self.pending = []

    def _write_layer(self, arr_hwc, path):
        np.save(path, np.ascontiguousarray(np.transpose(arr_hwc, (2, 0, 1))))          # (C, h, w) like v.cpu().numpy()[j]

    def _write_aggregate(self, maps_hwc, path):
        size = max(m.shape[1] for m in maps_hwc)                                       # resize_target = max width (:117-120)
        chw = [nearest_resize_chw(np.transpose(m, (2, 0, 1)), size) for m in maps_hwc]
        np.save(path, np.concatenate(chw, axis=0))

    def submit_batch(self, host_maps, first_index, count, done_event=None):
        """host_maps: {id: np.ndarray (B, h, w, C) fp16} already valid on the host."""
        a = self.args
        for j in range(count):
            name = sample_name(a, self.dataset, first_index + j)
            if a.aggregate_output:
                self.pending.append(self.pool.submit(self._write_aggregate, [m[j] for m in host_maps.values()],
                                                     output_path(a, name)))
            else:
                for k, m in host_maps.items():
                    self.pending.append(self.pool.submit(self._write_layer, m[j], output_path(a, name, k)))

    def drain(self):
        for f in self.pending:
            f.result()
        self.pending = []

    def close(self):
        self.drain()
        self.pool.shutdown()


class HostStager:
    """Asynchronous device-to-host copies of the captured maps: double-buffered pinned memory, a side stream, one
    event per batch. `stage` returns immediately; `wait` hands out numpy views once the copy has landed."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.turn = 0

    def stage(self, feats):
        s = self.turn
        self.turn ^= 1
        total = sum(v.numel() for v in feats.values())
        buf = self.slots[s]
        if buf is None or buf.numel() < total:
            buf = torch.empty(total, dtype=torch.float16).pin_memory()
            self.slots[s] = buf
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        views, off = {}, 0
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for k, v in feats.items():
                src = v.permute(0, 2, 3, 1)                   # the arena storage (B, h, w, C), contiguous
                dst = buf[off:off + src.numel()].view(src.shape)
                dst.copy_(src, non_blocking=True)
                views[k] = dst
                off += src.numel()
            done = torch.cuda.Event()
            done.record(self.stream)
        return views, done, list(feats.values())               # keep the device tensors alive until the copy is done

    @staticmethod
    def wait(staged):
        views, done, _keep = staged
        done.synchronize()
        return {k: v.numpy() for k, v in views.items()}


def prefetch_images(dataset, batch_size, img_size, depth=2, start=0, stop=None):
    """Generator of (first_index, [PIL images resized to img_size]) decoded by a background thread, over the
    [start, stop) shard of the dataset (indices stay global: they name the output files)."""
    from PIL import Image
    q = queue.Queue(maxsize=depth)
    stop = len(dataset) if stop is None else stop

    def work():
        try:
            for i in range(start, stop, batch_size):
                imgs = [Image.open(dataset[j][0]).resize((img_size, img_size)).convert("RGB")
                        for j in range(i, min(i + batch_size, stop))]
                q.put((i, imgs))
            q.put(None)
        except BaseException as ex:  # noqa: BLE001 - a corrupt / missing file must reach the consumer, not hang it
            q.put(ex)
    threading.Thread(target=work, daemon=True).start()
    while True:
        item = q.get()
        if item is None:
            return
        if isinstance(item, BaseException):
            raise RuntimeError("image prefetch failed: %r" % (item,)) from item
        yield item


def load_prompt_embeds(path):
    """The tuple `extract` takes as `prompts` (diffusion_feature.py:262-283), from a .pt (torch.save of the tuple
    encode_prompt returned, or of a dict) or a .safetensors file with the keys prompt_embeds [, negative_prompt_embeds,
    pooled_prompt_embeds, negative_pooled_prompt_embeds | prompt_attention_mask, negative_prompt_attention_mask]."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        d = load_file(path)
    else:
        d = torch.load(path, map_location="cpu")
    if isinstance(d, (tuple, list)):
        return tuple(d)
    pe = d["prompt_embeds"]
    if "prompt_attention_mask" in d:                       # PixArt
        return (pe, d["prompt_attention_mask"], d.get("negative_prompt_embeds", pe),
                d.get("negative_prompt_attention_mask", d["prompt_attention_mask"]))
    if "pooled_prompt_embeds" in d and pe.shape[-1] == 4096:    # Flux: (T5 sequence, CLIP pooled)
        return (pe, d["pooled_prompt_embeds"])
    return (pe, d.get("negative_prompt_embeds", pe), d.get("pooled_prompt_embeds"),
            d.get("negative_pooled_prompt_embeds", d.get("pooled_prompt_embeds")))


def shard_of_this_rank(n_items):
    """Multi-GPU runs (one process per GPU, e.g. torchrun): every rank extracts its own contiguous shard of the input
    list into the shared output tree - images are independent, so there is no collective (parallel.shard_range)."""
    from .parallel import shard_range
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    return shard_range(n_items, rank, world) if world > 1 else (0, n_items)


def run(args, extractor=None):
    from .diffusion_feature import FeatureExtractor
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.device == 'cuda':
        args.device = 'cuda:%d' % int(os.environ.get("LOCAL_RANK", "0"))
    os.makedirs(args.output_dir, exist_ok=True)
    print(f'Run folder: {args.output_dir}')
    if args.show_all_layers:
        args.layer = None
    if args.offline_lora or args.control:
        raise NotImplementedError("--offline_lora / --control are not built on the B200 path (SURVEY.md 2.1)")
    if extractor is None:
        # weights: --checkpoint, else $GDF_MODEL_DIR, else an explicit --synthetic; never a silent random network
        if getattr(args, "checkpoint", None):
            os.environ["GDF_MODEL_DIR"] = args.checkpoint
        elif getattr(args, "synthetic", False) and not os.environ.get("GDF_MODEL_DIR"):
            os.environ["GDF_SYNTHETIC"] = "1"
        elif not os.environ.get("GDF_MODEL_DIR"):
            raise SystemExit("no weights: pass --checkpoint <diffusers model dir> (or set GDF_MODEL_DIR), or opt in to "
                             "random-init weights with --synthetic")
    df = extractor or FeatureExtractor(args.layer, args.version, device=args.device, dtype=args.dtype,
                                       offline_lora=args.offline_lora, offline_lora_filename=args.offline_lora_filename,
                                       feature_resize=args.feature_resize, control=args.control,
                                       attention=args.attention, img_size=args.img_size)
    dataset = list_inputs(args.input_dir, args.nested_input_dir)
    with open(args.prompt_file, 'r') as f:
        prompts = f.read()
        print('prompt:', prompts)
    if getattr(args, "prompt_embeds", None):
        prompts = load_prompt_embeds(args.prompt_embeds)
    elif getattr(args, "synthetic", False) or extractor is not None:
        prompts = df.encode_prompt(prompts)   # stand-in embeddings seeded by the prompt text (warns)
    else:
        raise SystemExit("no prompt embeddings: text encoders are outside this path - pass --prompt_embeds <file> "
                         "(encode the prompt once with the reference's FeatureExtractor.encode_prompt and torch.save "
                         "the tuple), or opt in to stand-in embeddings with --synthetic")
    writer = NpyWriter(args, dataset, args.writer_threads)
    stager = None
    in_flight = None                           # (staged copy, first index, count) of the previous batch
    n_done = 0
    with torch.no_grad():
        lo, hi = shard_of_this_rank(len(dataset))
        for first, imgs in prefetch_images(dataset, args.batch_size, args.img_size, start=lo, stop=hi):
            features = df.extract(prompts, len(imgs), imgs, t=args.t, denoising_from=args.denoising_from,
                                  use_control=args.control is not None, use_ddim_inversion=args.use_ddim_inversion)
            if args.show_all_layers:             # debug mode of the reference (:100-108)
                layer_record = {}
                for k, v in features.items():
                    print(k, v[0].shape)
                    layer_record[k] = True
                with open('layer_record.json', 'w') as f:
                    f.write(json.dumps(layer_record))
                return 0
            if not next(iter(features.values())).is_cuda:   # accept_all returns host tensors (feature_extractor.py:65-66)
                host = {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in features.items()}
                writer.submit_batch(host, first, len(imgs))
                writer.drain()
            else:
                if stager is None:
                    stager = HostStager(next(iter(features.values())).device)
                staged = stager.stage(features)
                if in_flight is not None:        # write batch i-1 while the GPU runs batch i and copies it out
                    writer.submit_batch(HostStager.wait(in_flight[0]), in_flight[1], in_flight[2])
                    writer.drain()               # its pinned buffer is reused two batches later
                in_flight = (staged, first, len(imgs))
            n_done += len(imgs)
        if in_flight is not None:
            writer.submit_batch(HostStager.wait(in_flight[0]), in_flight[1], in_flight[2])
    writer.close()
    print('extracted %d images' % n_done)
    return n_done


def main(argv=None):
    return run(build_parser().parse_args(argv))


if __name__ == '__main__':
    main()
