"""A handful of launches of each hot kernel at the benchmark shapes, for `ncu --set full`."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import ops, raster as R, synthetic as S

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda"
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).to(dev).half()
if which in ("all", "attn"):
    T, N, H, D = 24, 512, 16, 32
    q = rn(T, N, H, D)
    kv = rn(4096, 2, H, D)
    kvi = rn(T, 1370, 2, H, D)
    qkv = rn(T, N, 3, H, D)
    for _ in range(4):
        ops.attention(q, kv[:, 0], kv[:, 1], 1 / math.sqrt(D), kv_shared=True)       # static cross
        ops.attention(q, kvi[:, :, 0], kvi[:, :, 1], 1 / math.sqrt(D))                # image cross
        ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 1 / math.sqrt(D))     # spatial self
    q64 = rn(8192, 12, 64)
    kv64 = rn(T, 512, 2, 12, 64)
    for _ in range(2):
        ops.attention(q64, kv64[:, :, 0], kv64[:, :, 1], 0.125, q_shared=True)        # VAE decoder cross
if which in ("all", "gemm"):
    a = rn(12288, 512)
    w1, w2 = rn(2048, 512), rn(512, 2048)
    b1 = torch.randn(2048, generator=g).to(dev)
    x = torch.randn(12288, 512, generator=g).to(dev)
    for _ in range(4):
        h = ops.gemm(a, w1, b1, ops.EPI_GELU_F16)
        ops.gemm(h, w2, None, ops.EPI_RESID_F32, out=x)
if which in ("all", "raster"):
    canon = S.canonical_gaussians(num_voxels=2048)
    P = canon["_xyz"].shape[0]
    delta = S.raster_delta(24, P).to(dev)
    cams, tfx, tfy = R.pack_cameras(S.orbit_extrinsics(24), S.intrinsics(), 0.8, 1.6)
    prm = R.make_params(512, 512, tfx, tfy, S.gaussian_constants())
    rz = R.Rasterizer(dev)
    arrays = R.canon_arrays(canon, dev)
    for _ in range(3):
        rz.forward(prm, arrays, delta, cams.to(dev), want_radii=False)
torch.cuda.synchronize()
print("done")
