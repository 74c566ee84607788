#!/usr/bin/env python
"""PCIe ceiling of the box for the end-to-end number: pinned host <-> device copies, one direction alone and both at once
(two streams), 1 GiB per direction, CUDA events."""
import json
import torch


def main():
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        b.record()
        torch.cuda.synchronize()
        return n * reps / (a.elapsed_time(b) * 1e-3) / 1e9

    run(True, True, 1)
    print(json.dumps({"h2d_alone_gbs": round(run(True, False), 1), "d2h_alone_gbs": round(run(False, True), 1),
                      "both_each_direction_gbs": round(run(True, True), 1)}))


if __name__ == "__main__":
    main()
