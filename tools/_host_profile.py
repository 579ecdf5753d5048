import cProfile, pstats, sys, time, torch
sys.path.insert(0, '.')
import tools.static_vae_step_bench as B
S = B.build(torch.device('cuda', 0))
for _ in range(3):
    B.step_ours(S)
torch.cuda.synchronize()
def fwd_only():
    t0 = time.perf_counter()
    terms, _ = S["fw"].training_losses(S["x"], S["image"], S["ext"], S["intr"], noise=S["noise"])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return terms, t1 - t0, t2 - t0
for _ in range(3):
    terms, a, b = fwd_only()
    t0 = time.perf_counter(); (terms["loss"] * 65536.0).backward(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"fwd enqueue {a*1e3:.2f} done {b*1e3:.2f} | bwd enqueue {(t1-t0)*1e3:.2f} done {(t2-t0)*1e3:.2f}")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    B.step_ours(S)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
