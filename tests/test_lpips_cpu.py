"""LPIPS-VGG16 criterion of the training step (gvfdiffusion_b200/utils/lpips) against the functional restatement in
oracle/lpips.py, on the CPU (plain torch modules: SURVEY.md row a17 keeps this term on library convolutions)."""
import torch


def test_lpips_matches_oracle_and_only_the_prediction_gets_gradients():
    from gvfdiffusion_b200.utils.lpips import LPIPS
    from oracle import lpips as OL
    m = LPIPS(net_type="vgg", seed=3).eval()
    assert not any(p.requires_grad for p in m.parameters())
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).requires_grad_(True)
    y = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    loss = m(x, y)
    ref = OL.lpips(m.layers.state_dict(), m.lin.state_dict(), x.detach(), y)
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref)), (float(loss), float(ref))
    loss.backward()
    assert x.grad is not None and float(x.grad.abs().sum()) > 0
    assert float(m(y, y)) == 0.0
    # state-dict layouts: torchvision's `features.N.*` and the published heads `linN.model.1.weight`
    vsd = {"features." + k: v.clone() for k, v in m.layers.state_dict().items()}
    lsd = {f"lin{i}.model.1.weight": l[1].weight.clone() * 2 for i, l in enumerate(m.lin)}
    m2 = LPIPS(vgg_state_dict=vsd, lin_state_dict=lsd).eval()
    assert m2.pretrained and abs(float(m2(x.detach(), y)) - 2 * float(ref)) < 1e-4 * abs(float(ref))


def test_lpips_matches_the_reference_class_fixture():
    """tests/golden/lpips.pt: loss and input gradient of the REFERENCE's own LPIPS(net_type='vgg') class run with the seeded
    random weights of LPIPS(seed=3) in place of its two downloads (tests/golden/make_golden.py gen_lpips)."""
    import os
    from gvfdiffusion_b200.utils.lpips import LPIPS
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lpips.pt"), weights_only=False)
    m = LPIPS(seed=fx["seed"]).eval()
    g = torch.Generator().manual_seed(fx["input_seed"])
    x = (torch.rand(*fx["shape"], generator=g) * 2 - 1).requires_grad_(True)
    y = torch.rand(*fx["shape"], generator=g) * 2 - 1
    loss = m(x, y)
    loss.backward()
    assert abs(float(loss) - fx["loss"]) < 1e-5 * abs(fx["loss"]), (float(loss), fx["loss"])
    assert abs(float(x.grad.abs().sum()) - fx["grad_abs_sum"]) < 1e-4 * fx["grad_abs_sum"]
    assert torch.allclose(x.grad[0, :, ::16, ::16], fx["grad_probe"], rtol=1e-3, atol=1e-9)
