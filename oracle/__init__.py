"""CPU oracle for the GVFDiffusion sampling + rendering hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (torch fp32 / numpy /
plain C) of the reference's algorithm for the path SURVEY.md section 8 scopes:

    dpm.py      NoiseScheduleVP, model_wrapper (v-pred, 3-way CFG), DPM-Solver++(2M),
                adaptive DPM-Solver-12, IDDPM p_sample         (model/dpmsolver.py,
                model/gaussian_diffusion.py, model/respace.py)
    dit.py      DiT._forward and its blocks                    (model/dit.py,
                model/attention/modules.py, model/attention/full_attn.py)
    vae.py      GSKLTemporalVariationalAutoEncoder.decode      (model/autoencoder.py)
    gaussian.py GaussianModel activations / get_*_with_delta, camera set-up
                (representations/gaussian/gaussian_model.py, renderers/gaussian_render.py)
    losses.py   ssim / L1 / KNN / interpolation loss of the training step (utils/loss_util.py,
                train_vae.py:486-586; pytorch3d.knn_points restated from its documented semantics)
    sparse_window.py  calc_window_partition, windowed sparse attention, swin SparseTransformerBlocks and the
                SparseTransformerVAE encode / decode trunks (sparse/attention/windowed_attn.py,
                model/sparse_voxel_diffusion/sparse_transformer*.py)
    sparse_vae.py     SparseVAE.to_representation; submanifold sparse convolution (third-party spconv, absent:
                restated from its published semantics, parity unpinned)
    vox2seq.py  voxel <-> sequence codes (pinned to the reference's own PyTorch twin)
    raster.c    tile rasteriser forward + backward restatement (third-party
                diff_gaussian_rasterization, mip-splatting fork -- NOT in /root/reference,
                un-pinned git HEAD in setup.sh:220-224: "parity unpinned", see raster.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  The product package (gvfdiffusion_b200) never does.

Pinning: dpm.py / dit.py / vae.py / gaussian.py / losses.py / sparse_window.py / sparse_vae.py (to_representation) are checked against the reference's
own Python imported from /root/reference in the build container
(tests/golden/make_golden.py writes the fixtures, tests/test_oracle_golden.py
checks them everywhere; tests/test_sparse_vae_cpu.py, tests/test_render_call_cpu.py likewise).  raster.c has no
reference-side golden for its arithmetic: parity unpinned -- its BOUNDARY (what the reference hands to the
third-party rasteriser) is pinned by tests/golden/render_call.pt.
"""
