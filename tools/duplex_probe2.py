#!/usr/bin/env python3
"""GPU box probe: the host-pointer pipeline's copy pattern (many ~1 MB copies, 3 streams + events) in torch."""
import time, torch
CH, NCH, SLOTS = 262144, 8, 4
hin = [torch.empty(CH * NCH, dtype=torch.float32).pin_memory() for _ in range(12)]
hout = [torch.empty(CH * NCH, dtype=torch.float32).pin_memory() for _ in range(6)]
din = [[torch.empty(CH, dtype=torch.float32, device="cuda") for _ in range(12)] for _ in range(SLOTS)]
s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def run(use_events=True, do_out=True, do_in=True, kernel=True):
    ev_in = [torch.cuda.Event() for _ in range(SLOTS)]; ev_cmp = [torch.cuda.Event() for _ in range(SLOTS)]; ev_out = [torch.cuda.Event() for _ in range(SLOTS)]
    used = [False] * SLOTS
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for c in range(NCH):
        k = c % SLOTS; sl = slice(c * CH, (c + 1) * CH)
        if do_in:
            with torch.cuda.stream(s_in):
                if used[k] and use_events: s_in.wait_event(ev_out[k])
                for a in range(12): din[k][a].copy_(hin[a][sl], non_blocking=True)
                ev_in[k].record(s_in)
        with torch.cuda.stream(s_cmp):
            if use_events: s_cmp.wait_event(ev_in[k])
            if kernel: din[k][6].mul_(1.0)
            ev_cmp[k].record(s_cmp)
        if do_out:
            with torch.cuda.stream(s_out):
                if use_events: s_out.wait_event(ev_cmp[k])
                for a in range(6): hout[a][sl].copy_(din[k][6 + a], non_blocking=True)
                ev_out[k].record(s_out)
        used[k] = True
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for name, kw in (("in only", dict(do_out=False)), ("out only", dict(do_in=False)), ("in+out, events", dict()), ("in+out, no events", dict(use_events=False))):
    run(**kw); print(f"{name}: {min(run(**kw) for _ in range(5)):.2f} ms")
