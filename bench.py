#!/usr/bin/env python
"""bench.py — headline measurement of the B200-native HyperNeRF per-ray hot path.

Workload (BASELINE.json configs[1]): HyperNeRF translation warp + bendy sheet (hyper_dim 2), 64 coarse + 64 fine
samples, ONE training step = forward + MSE loss + backward over a 65 536-ray batch of synthetic LLFF-shaped rays
(strong scaling: the global batch is fixed and sharded over N GPUs, one NCCL all-reduce of the flat 5.9 MB gradient
per step), plus the Adam update.  Metric: train rays/s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference        # the reference algorithm (oracle port, fp32 torch) on the host CPU cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GLOBAL_RAYS = 65536
N_COARSE, N_FINE = 64, 64
FWD_FLOP_PER_EVAL = 2 * 801536            # 801 536 MAC per sample evaluation (unpadded)
TRUNK_FLOP_PER_EVAL = 2 * 673664          # the same without TranslationField (100 480) and HyperSheetMLP (27 392)
EVALS_PER_RAY = N_COARSE + (N_COARSE + N_FINE)
METRIC = "train rays/s (fwd+bwd, 64+64 samples/ray, HyperNeRF warp+bendy-sheet)"
WORKLOAD = "cfg2: HyperNeRF translation warp + bendy_sheet (hyper_dim 2), 65536-ray train batch, 64 coarse + 64 fine"


# --------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference algorithm (oracle restatement, fp32 torch) on the host cores, bounded sample
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(n_rays=1024, repeats=2, warmup=1):
    from hypernerf_torch_b200 import synthetic
    from oracle import hypernerf_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic.make_state_dict(synthetic.cfg1_state_dict_shapes(), seed=0, boosted=False)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rays, rgbs = synthetic.train_rays(n_rays, seed=0)
    cfg = orc.default_cfg(n_fine=N_FINE, noise_std=1.0)
    best = float("inf")
    for it in range(warmup + repeats):
        g = torch.Generator().manual_seed(it)
        draws = dict(u_coarse=torch.rand(n_rays, N_COARSE, generator=g),
                     noise_coarse=torch.randn(n_rays, N_COARSE, 1, generator=g),
                     u_fine=torch.rand(n_rays, N_FINE, generator=g),
                     noise_fine=torch.randn(n_rays, N_COARSE + N_FINE, 1, generator=g))
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        out = orc.forward(sd, rays[:, :3], rays[:, 3:6], rays[:, 8].long(), draws, cfg)
        orc.mse_loss(out, rgbs).backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            best = min(best, dt)
    return n_rays / best, cores, f"{n_rays} rays of the cfg2 workload (64+64 samples, fwd+bwd, fp32 torch on CPU), best of {repeats}"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    rate, cores, sample = cpu_reference_rate(n_rays=1024, repeats=steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * 1024 / rate, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU arm: each step is a 1024-ray sample of the workload"},
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an NVML polling thread (first sample at once, then every
    20 ms, so even the 65 ms timed region of the 8-GPU run is covered); nvidia-smi -lms as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    MASKS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, gpu_index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            gpu_index = int(vis.split(",")[gpu_index]) if vis else gpu_index
        except (ValueError, IndexError):
            pass
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, threading.Event(), None

    def _poll(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while True:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                if reasons_fn is not None:
                    bits = int(reasons_fn(self.handle))
                    for name, mask in self.MASKS:
                        if bits & mask:
                            self.reasons.add(name)
            except Exception:
                pass
            if self.stop_flag.wait(0.02):
                return

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch.distributed as dist
    from hypernerf_torch_b200 import _lib, synthetic
    from hypernerf_torch_b200 import train as hn_train
    from hypernerf_torch_b200.models import NerfModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()  # fail loudly if the native library is missing

    emb = {'warp': list(range(100)), 'camera': [0], 'appearance': list(range(100)), 'time': list(range(100))}
    model = NerfModel(emb, near=0., far=1., n_samples_coarse=N_COARSE, n_samples_fine=N_FINE, noise_std=1.0,
                      hyper_slice_method='bendy_sheet', hyper_slice_out_dim=2, use_warp=True, use_nerf_embed=False,
                      use_alpha_cond=False, use_rgb_cond=False, GLO_dim=8, share_GLO=True, xyz_fourier_dim=10,
                      hyper_fourier_dim=6, view_fourier_dim=6)
    model.load_state_dict(synthetic.make_state_dict(model, seed=0, boosted=False))
    model = model.to(dev)
    fg = hn_train.FlatGrads(model.parameters())
    model.attach_flat_grads(fg)
    opt = hn_train.FusedAdam(fg, lr=5e-4, eps=1e-8)   # Adam of utils/__init__.py:29-31 / opt.py:53-56, one launch

    # inputs: this rank's contiguous shard of the global batch, resident in HBM (value) and in pinned host memory (e2e)
    rays_all, rgbs_all = synthetic.train_rays(GLOBAL_RAYS, seed=0)
    lo, hi = hn_train.shard_bounds(GLOBAL_RAYS, rank, world)
    rays_h = rays_all[lo:hi].contiguous().pin_memory()
    rgbs_h = rgbs_all[lo:hi].contiguous().pin_memory()
    rays_d, rgbs_d = rays_h.to(dev), rgbs_h.to(dev)
    chunk = args.chunk
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(rays, rgbs):
        return hn_train.train_step(model, rays, rgbs, fg, global_rays=GLOBAL_RAYS, chunk=chunk, optimizer=opt)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    torch.manual_seed(1234 + rank)
    for _ in range(max(args.warmup, 3)):
        step(rays_d, rgbs_d)
    sync_all()

    def timed_loop(fn, steps, profile=False):
        """EXACTLY `steps` steps; device time by CUDA events on the launching stream, L2 flushed between steps
        (outside the events); returns (seconds summed over steps, launches)."""
        evs = []
        launches0 = _lib.launches
        if profile:
            _lib.profile = []
        sync_all()
        for _ in range(steps):
            l2_flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        sync_all()
        secs = sum(a.elapsed_time(b) for a, b in evs) / 1e3
        prof, _lib.profile = _lib.profile, None
        t = torch.tensor([secs], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.launches - launches0, prof

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    secs, launches, prof = timed_loop(lambda: step(rays_d, rgbs_d), args.steps, profile=True)
    clk = clocks.stop() if rank == 0 else None

    # end to end: host rows in, loss out, every step
    loss_h = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        r = rays_h.to(dev, non_blocking=True)
        c = rgbs_h.to(dev, non_blocking=True)
        loss = step(r, c)
        loss_h.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    secs_e2e, _, _ = timed_loop(e2e_step, args.steps)

    # secondary metric of BASELINE.json ("render Mrays/s", configs[2]): one 1008x756 frame, 64+128 samples, no_grad,
    # random-init weights, rows of the frame sharded over the ranks (no collective); device time, max over ranks
    render = None
    if not args.no_render:
        rmodel = NerfModel(emb, near=0., far=1., n_samples_coarse=64, n_samples_fine=128, noise_std=None,
                           hyper_slice_method='bendy_sheet', hyper_slice_out_dim=2, use_warp=True, use_nerf_embed=False,
                           use_alpha_cond=False, use_rgb_cond=False, GLO_dim=8, share_GLO=True, xyz_fourier_dim=10,
                           hyper_fourier_dim=6, view_fourier_dim=6)
        rmodel.load_state_dict(model.state_dict())
        rmodel = rmodel.to(dev).eval()
        frame = synthetic.frame_rays(image_id=3, seed=0)
        flo, fhi = hn_train.shard_bounds(frame.shape[0], rank, world)
        frame_d = frame[flo:fhi].contiguous().to(dev)
        hn_train.render_rays(rmodel, frame_d[:32768], chunk=32768)       # warm-up (packs the weights)
        secs_r, launches_r, _ = timed_loop(lambda: hn_train.render_rays(rmodel, frame_d, chunk=32768), 2)
        render = {"value": frame.shape[0] * 2 / secs_r / 1e6, "unit": "Mrays/s", "rays_per_frame": int(frame.shape[0]),
                  "samples": "64+128", "ms_per_frame": 1e3 * secs_r / 2, "frames_timed": 2,
                  "workload": "cfg3: full-frame eval render 1008x756, random-init weights, no_grad"}
        del rmodel, frame_d

    # secondary workload (BASELINE.json configs[3]): static NeRF baseline (models/nerf.py x2 + render_rays), 262 144-ray
    # batch sharded over the ranks, 128+128 samples, perturb=1, noise_std=1, forward + backward (no optimizer)
    static = None
    if not args.no_static:
        from hypernerf_torch_b200.nerf import Embedding, NeRF
        from hypernerf_torch_b200.rendering import render_rays as static_render_rays
        smodels = [NeRF().to(dev), NeRF().to(dev)]
        semb = [Embedding(3, 10), Embedding(3, 4)]
        n_static = 262144
        slo, shi = hn_train.shard_bounds(n_static, rank, world)
        srays9, srgbs = synthetic.train_rays(shi - slo, seed=7 + rank, device=dev)
        srays = srays9[:, :8].contiguous()

        def static_step():
            for m in smodels:
                m.zero_grad(set_to_none=True)
            for i in range(0, srays.shape[0], 32768):
                out = static_render_rays(smodels, semb, srays[i:i + 32768], N_samples=128, perturb=1.0, noise_std=1.0,
                                         N_importance=128)
                tgt = srgbs[i:i + 32768]
                loss = (torch.nn.functional.mse_loss(out['rgb_coarse'], tgt, reduction='sum') +
                        torch.nn.functional.mse_loss(out['rgb_fine'], tgt, reduction='sum')) / (3.0 * n_static)
                loss.backward()

        static_step()
        secs_s, _, _ = timed_loop(static_step, 2)
        static = {"value": n_static * 2 / secs_s, "unit": "rays/s", "rays_per_step": n_static, "samples": "128+128",
                  "ms_per_step": 1e3 * secs_s / 2, "steps_timed": 2,
                  "workload": "cfg4: static NeRF baseline, 262144-ray batch, perturb=1, noise_std=1, fwd+bwd"}
        del smodels, srays, srgbs

    # roofline of the dominant kernel from the per-kernel events recorded inside the timed region
    per = {}
    for name, n, a, b in prof or []:
        d = per.setdefault(name, [0.0, 0, 0])
        d[0] += a.elapsed_time(b) / 1e3; d[1] += 1; d[2] += n
    roof, kernels = None, {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_hbm = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json (bf16_tflops_sustained / hbm_gbs: kernels timed inside a long step)" if peaks else \
        "fallback 1.4 PFLOP/s sustained, 6.65 TB/s (B200_PROFILING.md)"
    # Which roofline bounds each kernel: the fused forward and data-gradient kernels are dense
    # contractions (tensor pipe); the weight-gradient kernel streams both activation stashes once and is HBM bound.
    # Algorithmic HBM bytes per sample: X stash 8640 B + dY stash 8288 B read once by wgrad.
    # The fine level evaluates the depths it inherits from the coarse level with the trunk-only program (the shared warp /
    # sheet nets were already evaluated there by the coarse pass): 673 664 MAC per sample, and the stashes without the
    # warp / sheet slabs (X 6 176 B + dY 5 952 B).
    ALG_BYTES = {"mlp_wgrad": 8640 + 8288, "mlp_wgrad_trunk": 6176 + 5952}
    executed_flop = 0.0
    for name, (t, cnt, nsamp) in per.items():
        # fwd, dgrad and wgrad each count 1x forward FLOPs (bwd = 2x fwd)
        flops = (TRUNK_FLOP_PER_EVAL if name.endswith("_trunk") else FWD_FLOP_PER_EVAL) * nsamp
        executed_flop += flops
        kernels[name] = {"launches": cnt, "seconds": t, "tflops": flops / t / 1e12 if t > 0 else None,
                         "share_of_step": t / (secs if secs > 0 else 1)}
        if name in ALG_BYTES and t > 0:
            kernels[name]["hbm_gbs"] = ALG_BYTES[name] * nsamp / t / 1e9
    if kernels:
        dom = max(kernels, key=lambda k: kernels[k]["seconds"])
        k = kernels[dom]
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(dom)
            # ncu dram bytes per sample of this kernel x the average samples per launch of the timed region
            traffic = tr["dram_bytes_per_sample"] * per[dom][2] / per[dom][1]
        except Exception:
            pass
        if dom in ALG_BYTES:
            roof = {"kernel": dom, "bound": "hbm", "achieved": k["hbm_gbs"], "peak": peak_hbm, "unit": "GB/s",
                    "frac": k["hbm_gbs"] / peak_hbm, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ALG_BYTES[dom] * per[dom][2] / per[dom][1],
                    "avg_launch_ms": 1e3 * k["seconds"] / k["launches"], "kernels": kernels}
        else:
            roof = {"kernel": dom, "bound": "tensor", "achieved": k["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": k["tflops"] / peak_tf, "traffic": traffic, "peak_source": peak_src,
                    "avg_launch_ms": 1e3 * k["seconds"] / k["launches"], "kernels": kernels}
        # the whole step against the tensor roofline, for the north-star utilisation target
        # (executed FLOPs of this rank's launches: the redundant warp / sheet evaluations the reference makes at the
        # inherited depths are not counted)
        roof["step_tensor_frac"] = (executed_flop / secs / 1e12) / peak_tf
        roof["executed_mflop_per_ray"] = executed_flop / (args.steps * GLOBAL_RAYS / world) / 1e6
        roof["executed_tflops_per_gpu"] = executed_flop / secs / 1e12

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, sample = cpu_reference_rate(n_rays=1024, repeats=2, warmup=1)
            cpu = {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample}
        rays_per_s = GLOBAL_RAYS * args.steps / secs
        total_flop = GLOBAL_RAYS * EVALS_PER_RAY * FWD_FLOP_PER_EVAL * 3
        line = {
            "metric": METRIC, "value": rays_per_s, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_rays": GLOBAL_RAYS, "rays_per_gpu": hi - lo, "chunk_rays": chunk,
                       "samples": f"{N_COARSE}+{N_FINE}", "parallelism": f"dp{world} (ray shards, 1 flat NCCL all-reduce/step)",
                       "optimizer": "Adam step inside the timed region", "timing": "CUDA events per step, max over ranks",
                       "l2": "256 MB buffer written between timed steps; per-step working set (>10 GB) exceeds L2",
                       "flop_accounting": "model_tflops counts the reference's 923.4 MFLOP per ray (192 full-network "
                                          "evaluations); roofline.* count executed FLOPs (the fine level's 64 inherited "
                                          "depths skip the shared warp / sheet nets)"},
            "model_tflops": total_flop * args.steps / secs / 1e12,
            "e2e": {"value": GLOBAL_RAYS * args.steps / secs_e2e, "unit": "rays/s",
                    "h2d_bytes_per_step": int(rays_h.numel() * 4 + rgbs_h.numel() * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "render": render, "static_nerf": static,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract goes to the process's real stdout; everything else any library prints on
    fd 1 during the run (e.g. NCCL's version banner at N > 1) has been redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunk", type=int, default=8192, help="rays per forward/backward chunk on one GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip the secondary full-frame render measurement")
    ap.add_argument("--no-static", action="store_true", help="skip the secondary static-NeRF (cfg4) measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
