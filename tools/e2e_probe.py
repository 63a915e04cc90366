"""Where does MultiscaleTrainer.train() lose time against the device-timed step?  A/B of the host-side options on one
box (balloons, batch 32):   gpurun -- python tools/e2e_probe.py [steps]"""
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    job = bench.Job()
    torch.manual_seed(0)
    tr = bench.make_trainer(job, bench.BALLOONS_SIZES, bench.BALLOONS_T_IDEAL, bench.BATCH, "probe")
    for i in range(10):
        tr.train_step(s=i % 5)
    torch.cuda.synchronize()
    dev_data = tr.data_list
    host_data = [tuple(t.cpu().pin_memory() for t in pair) for pair in dev_data]

    def run(label, host, readback, draw, side=True, seq=None):
        tr.host_data = host
        tr.data_list = host_data if host else dev_data
        tr.loss_readback = readback
        tr.scale_draw = draw
        torch.manual_seed(235)
        with bench.stdout_to_stderr():
            tr._prepare_training()
            if not side:
                tr._draw_stream = None
            tr.scale_counts = [0] * 5
            draw_s = []
            if seq is not None:
                it = iter(seq)
                tr._draw_scale = lambda: next(it)
            else:
                orig = type(tr)._draw_scale.__get__(tr)

                def timed_draw():
                    t0 = time.perf_counter()
                    s = orig()
                    draw_s.append(time.perf_counter() - t0)
                    return s
                tr._draw_scale = timed_draw
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                tr.train_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            del tr._draw_scale
            if not side:
                del tr._draw_stream
        extra = f" draw host ms mean {1e3 * sum(draw_s) / len(draw_s):.3f} max {1e3 * max(draw_s):.3f}" if draw_s else ""
        print(f"{label:58s} {1e3 * dt / steps:8.3f} ms/step counts {tr.scale_counts}{extra}", flush=True)
        return list(tr.scale_counts)

    if len(sys.argv) > 2 and sys.argv[2] == "bench":
        # the end-to-end leg of bench.py with a timestamp per step
        import argparse
        tr.host_data = True
        tr.data_list = host_data
        tr.loss_readback = "step"
        seed, predicted = bench.pick_balanced_seed(tr, job, bench.BALLOONS_SIZES, steps, 5)
        print("seed", seed, "predicted counts", predicted, flush=True)
        torch.manual_seed(seed)
        stamps = []
        # phase timers: anything on the host that takes > 3 ms is reported with its name
        slow = []

        def timed(obj, name, label):
            fn = getattr(obj, name)

            def wrapper(*a, **k):
                t0 = time.perf_counter()
                out = fn(*a, **k)
                dt = time.perf_counter() - t0
                if dt > 3e-3:
                    slow.append((label, round(1e3 * dt, 2), tr.step))
                return out
            setattr(obj, name, wrapper)
        timed(tr, "_draw_scale", "draw_scale")
        timed(tr, "_batch", "batch_copy")
        timed(tr.model, "forward", "diffusion.forward(enqueue)")
        timed(tr._fused, "step", "fused.step(enqueue)")
        timed(tr.scheduler, "step", "scheduler.step")
        import sinddm_b200.trainer as T
        orig_bw = T.loss_backwards

        def bw(*a, **k):
            t0 = time.perf_counter()
            orig_bw(*a, **k)
            dt = time.perf_counter() - t0
            if dt > 3e-3:
                slow.append(("backward(enqueue)", round(1e3 * dt, 2), tr.step))
        T.loss_backwards = bw
        inner = tr.train_step

        def stamped():
            t0 = time.perf_counter()
            out = inner()
            stamps.append((time.perf_counter() - t0, tr.step))
            return out
        tr.train_step = stamped
        with bench.stdout_to_stderr():
            tr.train_num_steps = tr.step + 5
            tr.train()
            torch.cuda.synchronize()
            warm = len(stamps)
            tr.scale_counts = [0] * 5
            t0 = time.perf_counter()
            tr.train_num_steps = tr.step + steps
            tr.train()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"train() {1e3 * dt / steps:.3f} ms/step, counts {tr.scale_counts}", flush=True)
        print("per-step host ms:", " ".join(f"{1e3 * a:.1f}" for a, _ in stamps[warm:]), flush=True)
        print("slow host phases (> 3 ms; loss.item() waits are not wrapped):", slow, flush=True)
        return
    seq = [i % 5 for i in range(steps)]
    run("round robin, device data, window readback", False, "window", "device", seq=seq)
    run("round robin, HOST data, window readback", True, "window", "device", seq=seq)
    run("round robin, device data, loss.item() per step", False, "step", "device", seq=seq)
    run("round robin, HOST data, loss.item() per step", True, "step", "device", seq=seq)
    run("multinomial side stream, device data, window", False, "window", "device")
    run("multinomial IN stream, device data, window", False, "window", "device", side=False)
    run("multinomial host generator, device data, window", False, "window", "host")
    run("multinomial side stream, HOST data, item per step", True, "step", "device")
    run("multinomial side stream, device data, item per step", False, "step", "device")


if __name__ == "__main__":
    main()
