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

# the e2e pipeline's copy skeleton without any search: 4 shrinking chunks in, each chunk's results out as soon
# as its input has landed (weights 5,4,3,2) -- the floor of `e2e` set by PCIe alone
def skeleton():
    w = [5, 4, 3, 2]
    tot = sum(w)
    e = torch.cuda.Event(); e.record()
    s1.wait_event(e); s2.wait_event(e)
    off = 0
    for wi in w:
        m = n * wi // tot
        with torch.cuda.stream(s1):
            d_in[off:off + m].copy_(h_in[off:off + m], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(s1)
        s2.wait_event(ev)
        with torch.cuda.stream(s2):
            h_out[off:off + m].copy_(d_out[off:off + m], non_blocking=True)
        off += m
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
import time
for _ in range(3): skeleton()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    skeleton(); torch.cuda.synchronize()
print(json.dumps({"skeleton_4chunks_wall_ms": (time.perf_counter() - t0) / 20 * 1e3, "h2d_again_ms": t(h2d), "d2h_again_ms": t(d2h)}))
