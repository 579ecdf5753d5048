"""One-command pinning of the rasteriser oracle against the UPSTREAM extension.

`oracle/raster.c` restates `diff_gaussian_rasterization` (autonomousvision/mip-splatting,
submodules/diff-gaussian-rasterization, installed by the reference's setup.sh:220-224 at an un-pinned HEAD), which is
absent from /root/reference and from this image -- its header says "parity unpinned".  On any CUDA machine where
that extension IS importable:

    GVF_UPSTREAM_RASTER=/path/to/site-packages-or-checkout  python tests/golden/make_raster_golden.py

renders the `tests/_scenes.scene` fixtures through upstream with exactly the settings the reference's
`renderers/gaussian_render.py:105-125,198-206` passes (recorded in tests/golden/render_call.pt) and writes
`tests/golden/raster_upstream.pt` (colour [3,H,W], radii [P] per frame + the upstream commit if it can be read).
`tests/test_raster_upstream_pinning_cpu.py` then compares the oracle with it (skipped while the file is absent);
when it passes, delete the "PARITY UNPINNED" paragraph of oracle/raster.c and DESIGN.md section 6.
"""
import os
import subprocess
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SCENES = [dict(num_voxels=256, F=3, H=128, W=128, seed=0, with_delta=True, scale_boost=0.0),
          dict(num_voxels=128, F=2, H=96, W=160, seed=3, with_delta=True, scale_boost=1.5),   # large splats, W != H
          dict(num_voxels=512, F=2, H=512, W=512, seed=5, with_delta=False, scale_boost=0.0)]


def main():
    hint = os.environ.get("GVF_UPSTREAM_RASTER")
    if not hint:
        sys.exit("set GVF_UPSTREAM_RASTER to the directory that contains the upstream diff_gaussian_rasterization package")
    if os.path.isdir(hint):
        sys.path.insert(0, hint)
    import diff_gaussian_rasterization as U                 # the UPSTREAM extension, not gvfdiffusion_b200's shim
    if "gvfdiffusion_b200" in (getattr(U, "__file__", "") or ""):
        sys.exit("GVF_UPSTREAM_RASTER resolves to this repo's shim; point it at the upstream build")
    if not torch.cuda.is_available():
        sys.exit("the upstream extension is CUDA only")
    from oracle import gaussian as G, raster as OR
    from tests import _scenes
    commit = ""
    try:
        commit = subprocess.check_output(["git", "-C", os.path.dirname(U.__file__), "rev-parse", "HEAD"], text=True).strip()
    except Exception:
        pass
    out = {"upstream_file": U.__file__, "upstream_commit": commit, "scenes": []}
    dev = "cuda"
    for sc in SCENES:
        canon, delta, ext, intr, const = _scenes.scene(**sc)
        cn = {k: v.numpy() for k, v in canon.items()}
        frames = []
        for f in range(sc["F"]):
            vt, pt, campos, tfx, tfy = G.camera_matrices(ext[f], intr, 0.8, 1.6)
            prm = OR.make_params(sc["H"], sc["W"], tfx, tfy, const, kernel_size=0.1, bg=(1.0, 1.0, 1.0))
            m3, scl, rot, sh, op = OR.activate(prm, cn, None if delta is None else delta[f].numpy())
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            settings = U.GaussianRasterizationSettings(
                image_height=sc["H"], image_width=sc["W"], tanfovx=float(tfx), tanfovy=float(tfy), kernel_size=0.1,
                subpixel_offset=torch.zeros((sc["H"], sc["W"], 2), dtype=torch.float32, device=dev),
                bg=torch.ones(3, dtype=torch.float32, device=dev), scale_modifier=1.0, viewmatrix=vt.to(dev),
                projmatrix=pt.to(dev), sh_degree=0, campos=campos.to(dev), prefiltered=False, debug=False)
            means3D = t(m3)
            color, radii = U.GaussianRasterizer(raster_settings=settings)(
                means3D=means3D, means2D=torch.zeros_like(means3D), shs=t(sh).reshape(-1, 1, 3), colors_precomp=None,
                opacities=t(op).reshape(-1, 1), scales=t(scl), rotations=t(rot), cov3D_precomp=None)
            frames.append({"color": color.cpu(), "radii": radii.cpu().to(torch.int32)})
        out["scenes"].append({"args": sc, "frames": frames})
    torch.save(out, os.path.join(HERE, "raster_upstream.pt"))
    print("wrote tests/golden/raster_upstream.pt from", U.__file__, commit)


if __name__ == "__main__":
    main()
