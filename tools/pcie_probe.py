#!/usr/bin/env python3
"""Host<->device copy rates of this box (pinned memory), alone and both directions at once: the bound of
bench.py's `e2e` leg, which moves 5.1 GB in and 2.3 GB out per ResNet-50 step.  CSV on stdout."""
import torch


def main():
    dev = torch.device("cuda:0")
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    run(True, True, 1)
    print("case,ms_per_GiB,h2d_GBs,d2h_GBs")
    t = run(True, False)
    print(f"h2d alone,{t:.2f},{n / t / 1e6:.1f},")
    t = run(False, True)
    print(f"d2h alone,{t:.2f},,{n / t / 1e6:.1f}")
    t = run(True, True)
    print(f"both directions,{t:.2f},{n / t / 1e6:.1f},{n / t / 1e6:.1f}")


if __name__ == "__main__":
    main()
