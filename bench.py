#!/usr/bin/env python
"""bench.py -- clips/s of the GAST-Net lifting forward (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--clips B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): 27-frame 17-joint inference, 4096 clips per GPU
(weak scaling, clips are independent -> no collective on the data path), fp32,
SpatioTemporalModel([3,3,3], channels=128) in eval mode, synthetic keypoints, seeded weights.
One step = one forward over the GPU's whole 4096-clip batch.

Prints ONE JSON line (rank 0).  `value` times the hot path with inputs resident in HBM;
`e2e` times the same call through the drop-in module with pinned HOST buffers (H2D of the
clips and D2H of the 3D poses inside the timed region).  `--impl reference` times the CPU
port of the reference algorithm (oracle/) on the host cores instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, 'gast-net-3dposeestimation_b200')
for _p in (PKG, REPO):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = '27-frame 17-joint clips/sec'
UNIT = 'clips/s'
J, FW, CH, T = 17, [3, 3, 3], 128, 27
FLOP_PER_CLIP = 0.402e9          # algorithmic (needed-only) forward FLOPs, BASELINE.md §2
IO_BYTES_PER_CLIP = 3876         # compulsory (T*J*2 + J*3)*4 bytes


def measured_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = set()
        for r in self.rows:
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        # the busiest half of the samples = "under load"
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {'sm_mhz': statistics.median(load) if load else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def build_model(device):
    import torch
    from gast_b200 import synth
    from model.gast_net import SpatioTemporalModel
    from common.skeleton import Skeleton
    from common.graph_utils import adj_mx_from_skeleton
    adj = adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))
    m = SpatioTemporalModel(adj, J, 2, J, FW, causal=False, dropout=0.05, channels=CH)
    synth.randomize_module(m, 1)
    return m.to(device).eval()


_CPU_STATE = {}


def cpu_port_clips_per_s(sample_clips, reps, threads=None):
    """Times the CPU port of the reference (oracle/gast_torch_ref.py: the same torch CPU kernels in
    the same order as the reference modules -- bit-identical outputs, same cost) on all host
    cores.  Needed-only (Optimized1f) schedule = the reference's fastest CPU form."""
    import torch
    from oracle import gast_oracle as O
    from oracle import gast_torch_ref as TR
    from gast_b200 import synth
    if 'p' not in _CPU_STATE:
        from model.gast_net import SpatioTemporalModelOptimized1f
        from common.skeleton import Skeleton
        from common.graph_utils import adj_mx_from_skeleton
        adj_t = adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))
        m = SpatioTemporalModelOptimized1f(adj_t, J, 2, J, FW, channels=CH)
        synth.randomize_module(m, 1)
        _CPU_STATE['p'] = {k: v.clone() for k, v in m.state_dict().items()}
        _CPU_STATE['masks'] = tuple(torch.from_numpy(a) for a in
                                    O.local_masks(O.adj_from_parents(synth.skeleton_parents(J))))
        # give the CPU baseline its best thread count: all host threads are not always the
        # fastest for these small ATen kernels
        xw = torch.from_numpy(synth.synth_input(16, T, J, 2, seed=1))
        best = (None, 1e30)
        cands = [threads] if threads else sorted({os.cpu_count(), 32, 16, 8} & set(range(1, os.cpu_count() + 1)), reverse=True)
        with torch.no_grad():
            for nt in cands:
                torch.set_num_threads(nt)
                TR.forward(xw, _CPU_STATE['p'], _CPU_STATE['masks'], FW, strided=True)      # warm
                t0 = time.perf_counter()
                TR.forward(xw, _CPU_STATE['p'], _CPU_STATE['masks'], FW, strided=True)
                dt = time.perf_counter() - t0
                if dt < best[1]:
                    best = (nt, dt)
        _CPU_STATE['threads'] = best[0]
    torch.set_num_threads(_CPU_STATE['threads'])
    x = torch.from_numpy(synth.synth_input(sample_clips, T, J, 2, seed=1234))
    ts = []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            TR.forward(x, _CPU_STATE['p'], _CPU_STATE['masks'], FW, strided=True)
            ts.append(time.perf_counter() - t0)
    return sample_clips / statistics.median(ts), torch.get_num_threads(), ts


def run_reference(args, rank, world):
    """`--impl reference`: the reference algorithm's CPU port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    sample = args.cpu_clips
    vals = []
    total_steps = args.warmup + args.steps
    per_step = []
    for s in range(total_steps):
        v, cores, ts = cpu_port_clips_per_s(sample, 1)
        if s >= args.warmup:
            vals.append(v)
            per_step.append(ts[0])
    value = sample * len(per_step) / sum(per_step)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * sum(per_step) / len(per_step),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': '27f/17j/128ch SpatioTemporalModel forward, eval, fp32 (BASELINE configs[1])',
                   'clips_per_step': sample, 'frames': T, 'joints': J, 'channels': CH,
                   'impl': 'CPU port of the reference on the host cores'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d clips per step, needed-only (Optimized1f) schedule, torch-CPU port of the '
                                   'reference (oracle/gast_torch_ref.py, bit-identical to it)' % sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--clips', type=int, default=4096, help='clips per GPU per step')
    ap.add_argument('--cpu-clips', type=int, default=64, help='clips per CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback '
                         '(use --impl reference for the CPU port)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    from gast_b200 import synth

    model = build_model(device)
    B = args.clips
    K, W = args.steps, max(args.warmup, 3)
    # rotating input buffers: 10 x 15 MB > 126 MB L2, so no step re-reads a cached input;
    # the per-step intermediates (~2.6 GB) already stream far more than L2 holds.
    NBUF = 10
    xs_host = [torch.from_numpy(synth.synth_input(B, T, J, 2, seed=1234 + rank * 100 + i)).pin_memory()
               for i in range(NBUF)]
    xs_dev = [x.to(device) for x in xs_host]
    y_host = torch.empty((B, 1, J, 3), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---------------- device-resident throughput (`value`) -----------------------------
    with torch.no_grad():
        for i in range(W):
            model(xs_dev[i % NBUF])
        launches_per_step = int(model._gast_last_launches)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(K):
            y = model(xs_dev[i % NBUF])
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None

        # ---------------- end to end through the public API with host buffers ----------
        for i in range(3):
            y_host.copy_(model(xs_host[i % NBUF].to(device, non_blocking=True)), non_blocking=True)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for i in range(K):
            xd = xs_host[i % NBUF].to(device, non_blocking=True)     # H2D of this step's clips
            y_host.copy_(model(xd), non_blocking=True)                 # D2H of this step's poses
        f1.record(stream)
        barrier()
        ms_e2e_serial = f0.elapsed_time(f1)
        # the serving API for host-resident clips: upload of step i+1 and download of step i-1 overlap the
        # forward of step i (gast_b200/stream.py); every step's H2D and D2H is inside the timed region
        from gast_b200.stream import PipelinedLifter
        lifter = PipelinedLifter(model, depth=2)
        y_hosts = [torch.empty((B, 1, J, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        lifter.run([xs_host[i % NBUF] for i in range(3)], [y_hosts[i % 2] for i in range(3)])
        barrier()
        f0.record(stream)
        lifter.run([xs_host[i % NBUF] for i in range(K)], [y_hosts[i % 2] for i in range(K)])
        f1.record(stream)
        barrier()
        ms_e2e = f0.elapsed_time(f1)

        # ---------------- per-kernel timing for the roofline ----------------------------
        from gast_b200 import engine
        prof = engine.profile_forward(model, xs_dev[0], reps=max(3, min(K, 10)))

    t = torch.tensor([ms, ms_e2e, ms_e2e_serial], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_e2e_serial = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_clips = B * world * K
    value = total_clips / (ms / 1000.0)
    e2e_value = total_clips / (ms_e2e / 1000.0)
    peaks, peak_src = measured_peaks()
    # dominant kernel: the fused GEMM family (all launches of it in one step)
    gemm_ms = prof['gemm_ms_per_step']
    gemm_flops = FLOP_PER_CLIP * B      # >99% of the algorithmic FLOPs are the channel contractions
    achieved_tf = gemm_flops / (gemm_ms / 1000.0) / 1e12
    peak_tf = peaks.get('bf16_tflops_sustained', peaks['bf16_tflops'])
    traffic = None
    tp = os.path.join(REPO, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get('dram_bytes_per_step')
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': '27f/17j/128ch SpatioTemporalModel forward, eval, fp32 (BASELINE configs[1])',
                   'clips_per_gpu': B, 'global_clips': B * world, 'frames': T, 'joints': J, 'channels': CH,
                   'parallelism': 'clip-sharded x%d, no collective' % world,
                   'l2': 'rotating %d input buffers (%.0f MB > 126 MB L2); intermediates %.1f GB/step'
                         % (NBUF, NBUF * B * T * J * 2 * 4 / 1e6, prof['workspace_bytes'] / 1e9),
                   'gemm_core': prof['gemm_core']},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': B * T * J * 2 * 4,
                'd2h_bytes_per_step': B * J * 3 * 4, 'ms_per_step': ms_e2e / K,
                'api': 'gast_b200.stream.PipelinedLifter (copy stream overlaps H2D/D2H with the forward)',
                'serial_value': total_clips / (ms_e2e_serial / 1000.0)},
        'gpu_launches': launches_per_step * K,
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                     'frac': achieved_tf / peak_tf, 'frac_of_fp32_parity_bound': achieved_tf / (peak_tf / 6.0),
                     'traffic': traffic,
                     'kernel': 'fused GEMM family (%d launches/step)' % prof['gemm_launches'],
                     'kernel_ms_per_step': gemm_ms, 'kernel_share_of_step': gemm_ms / prof['step_ms'],
                     'peak_source': 'bf16_tflops_sustained of %s MEASURED_PEAKS; fp32-parity 3xTF32 bound is peak/6'
                                    % peak_src,
                     'hbm_frac_compulsory': IO_BYTES_PER_CLIP * value / world / 1e9 / peaks['hbm_gbs'],
                     'per_kernel_ms': prof['per_kernel_ms']},
    }
    if not args.no_cpu_baseline:
        v, cores, ts = cpu_port_clips_per_s(args.cpu_clips, 5)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': '%d clips x 5 reps (median), needed-only schedule, torch-CPU port of the '
                                          'reference (oracle/gast_torch_ref.py)' % args.cpu_clips}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
