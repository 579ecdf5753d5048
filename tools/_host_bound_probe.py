import sys, time, torch
sys.path.insert(0, '.')
import tools.static_vae_step_bench as B
S = B.build(torch.device('cuda', 0))
for _ in range(3):
    B.step_ours(S)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    loss, e = B.step_ours(S)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3*(t1-t0):.2f} ms, until done {1e3*(t2-t0):.2f} ms, gpu fwd {e[0].elapsed_time(e[1]):.2f} bwd {e[1].elapsed_time(e[2]):.2f}")
