"""A/B of this repo's attention kernels against flash_attn 2.8.x (the kernels the reference itself
runs on a B200: reference model/attention/full_attn.py:114-120, model/autoencoder.py:132-144) on
the four DiT shapes and the two motion-VAE shapes of BASELINE configs[1].

    python tools/attn_vs_fa2.py [--out profiles/r02_attn_vs_fa2.json]

Warm = back-to-back launches on resident inputs; cold = a 256 MB buffer is rewritten before every
launch (L2 flushed).  CUDA events on the current stream, median of `--iters`.  flash_attn is
LIBRARY code used here only as the measured opponent; nothing in the product imports it.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvfdiffusion_b200 import ops  # noqa: E402


def timed(fn, iters, flush=None):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for a, b in ev:
        if flush is not None:
            flush.add_(1.0)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    try:
        import flash_attn
        from flash_attn import flash_attn_func, flash_attn_kvpacked_func
        fa_ver = flash_attn.__version__
    except Exception as e:      # noqa: BLE001
        print(json.dumps({"unavailable": f"flash_attn import failed: {e}"}))
        return
    dev = "cuda:0"
    g = torch.Generator(device="cpu").manual_seed(0)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)

    def rnd(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32).to(dev, torch.float16)

    # (name, Nb, Lq, Lk, H, D, kv_shared)
    shapes = [
        ("dit_static_cross", 24, 512, 4096, 16, 32, True),
        ("dit_image_cross", 24, 512, 1370, 16, 32, False),
        ("dit_spatial_self", 24, 512, 512, 16, 32, False),
        ("dit_temporal_self", 512, 24, 24, 16, 32, False),
        ("vae_self", 24, 512, 512, 12, 64, False),
        ("vae_decoder_cross", 24, 8192, 512, 12, 64, False),
    ]
    rows = []
    for name, Nb, Lq, Lk, H, D, shared in shapes:
        q = rnd(Nb, Lq, H, D)
        scale = D ** -0.5
        if shared:
            kv1 = rnd(Lk, 2, H, D)
            k1, v1 = kv1[:, 0], kv1[:, 1]
            # the reference repeats the static embedding over T (model/dit.py:465) before to_kv
            kv_rep = kv1.unsqueeze(0).expand(Nb, -1, -1, -1, -1).contiguous()
            ours = lambda: ops.attention(q, k1, v1, scale, kv_shared=True)      # noqa: E731
            fa = lambda: flash_attn_kvpacked_func(q, kv_rep, softmax_scale=scale)  # noqa: E731
        elif Lq == Lk and D == 32:
            # DiT self-attention (model/attention/modules.py:113-130): fused to_qkv output [N, L, 3, H, d]; the
            # reference hands flash_attn_func the RMS-normed q, k (fresh contiguous tensors) and v as a view.
            # Ours reads the packed tensor in place; the temporal one as the strided (N, T) view of the
            # (T, N, 3, H, d) layout (the reference transposes to contiguous first; those copies are not timed).
            if name == "dit_temporal_self":
                qkv = rnd(Lq, Nb, 3, H, D).permute(1, 0, 2, 3, 4)          # [N, T, 3, H, d] strided view
            else:
                qkv = rnd(Nb, Lq, 3, H, D)
            qc, kc, vc = qkv[:, :, 0].contiguous(), qkv[:, :, 1].contiguous(), qkv[:, :, 2].contiguous()
            out_buf = torch.empty_like(qkv[:, :, 0])                        # same strided layout as the engine's AO view
            ours = lambda: ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale, out=out_buf)  # noqa: E731
            fa = lambda: flash_attn_func(qc, kc, vc, softmax_scale=scale)   # noqa: E731
        else:
            kv = rnd(Nb, Lk, 2, H, D)
            k, v = kv[:, :, 0], kv[:, :, 1]
            ours = lambda: ops.attention(q, k, v, scale)                         # noqa: E731
            fa = (lambda: flash_attn_kvpacked_func(q, kv, softmax_scale=scale)) if D == 32 else \
                 (lambda: flash_attn_func(q, k, v, softmax_scale=scale))         # noqa: E731 (VAE: autoencoder.py:132-144)
        o1, o2 = ours().float(), fa().float()
        err = float((o1 - o2).norm() / o2.norm())
        flops = 4.0 * Nb * H * Lq * Lk * D
        r = {"shape": name, "Nb": Nb, "Lq": Lq, "Lk": Lk, "H": H, "D": D, "gflop": flops / 1e9,
             "rel_l2_vs_fa2": err}
        for tag, fl in (("warm", None), ("cold", flush)):
            t_o, t_f = timed(ours, args.iters, fl), timed(fa, args.iters, fl)
            r[f"ours_us_{tag}"], r[f"fa2_us_{tag}"] = round(t_o, 2), round(t_f, 2)
            r[f"ours_tflops_{tag}"], r[f"fa2_tflops_{tag}"] = round(flops / t_o / 1e6, 1), round(flops / t_f / 1e6, 1)
            r[f"speedup_{tag}"] = round(t_f / t_o, 3)
        rows.append(r)
        print(json.dumps(r), flush=True)
    res = {"flash_attn": fa_ver, "gpu": torch.cuda.get_device_name(0), "iters": args.iters, "rows": rows,
           "note": "cold = 256 MB buffer rewritten before each launch (flush time excluded: events bracket the call only)"}
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
