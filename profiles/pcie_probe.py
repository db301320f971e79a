"""PCIe ceiling of the host call: H2D alone, D2H alone, both directions at once (pinned memory, one copy stream per
direction, chunked like saa_step_host_ex) — the denominator of the e2e number.  python profiles/pcie_probe.py [MB]"""
import json
import sys
import time

import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 836
n = mb * (1 << 20) // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, chunks, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for c in range(chunks):
            a, b = n * c // chunks, n * (c + 1) // chunks
            if up:
                with torch.cuda.stream(s_in):
                    d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            if down:
                with torch.cuda.stream(s_out):
                    h_out[a:b].copy_(d_out[a:b], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


out = {"mb_per_direction": mb}
for chunks in (1, 32):
    run(True, True, chunks, 2)
    t_up, t_dn, t_both = run(True, False, chunks), run(False, True, chunks), run(True, True, chunks)
    out[f"chunks_{chunks}"] = {"h2d_gbs": n * 8 / t_up / 1e9, "d2h_gbs": n * 8 / t_dn / 1e9,
                              "both_gbs_per_direction": n * 8 / t_both / 1e9, "both_ms": t_both * 1e3}
print(json.dumps(out))
