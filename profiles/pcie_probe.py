"""Host<->device copy rates on this box (pinned memory), alone and in both directions at once:
the floor of the `e2e` figure (16 MB of queries in, 16 MB of results out per step)."""
import torch, json
n = 16 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def h2d(): d_in.copy_(h_in, non_blocking=True)
def d2h(): h_out.copy_(d_out, non_blocking=True)
def both():
    e = torch.cuda.Event(); e.record()
    s1.wait_event(e); s2.wait_event(e)
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
r = {"bytes": n, "h2d_ms": t(h2d), "d2h_ms": t(d2h), "both_ms": t(both)}
r["h2d_GBps"] = n / r["h2d_ms"] / 1e6; r["d2h_GBps"] = n / r["d2h_ms"] / 1e6
print(json.dumps(r))
