#!/usr/bin/env python
"""bench.py -- env-steps/s of the CrossroadEnd2end model rollout (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one H=25 `EnvironmentModel.rollout_out` rollout over the batch (BASELINE config #3:
B=65536 rows per GPU, V=32 vehicles, n=0 -> D=137, task left, mode='training' with per-row
reference paths): 25 fused `k_model_step` launches replayed as one CUDA graph.  N>1 (torchrun,
one rank per GPU) shards rows over ranks with no data-path collective (weak scaling: every rank
owns B rows); `value` = all ranks' env-steps / max-over-ranks device time.

Printed JSON line (rank 0): the driver contract plus
  roofline      algorithmic bytes (8*D+32 per env-step) / mean launch duration vs measured HBM peak
  cpu_baseline  the NumPy oracle (the reference's algorithm restated; TF2 is not installable
                offline) timed on the host cores on a bounded sample, N=1 only
  e2e           the same metric through the public EnvironmentModel API from pinned HOST buffers
                (H2D of observations and each step's actions, D2H of each step's 5 outputs and the
                final observations inside the timed region)
`--impl reference` times the oracle port itself (all host cores) as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'env-steps/s (ego x horizon)'
TASK, V, N_FUT, H = 'left', 32, 0, 25
D = 6 + 3 * (N_FUT + 1) + 4 * V
BYTES_PER_ENV_STEP = 8 * D + 32          # obs in+out, action 8, ref_index 4, 5 outputs 20


def workload_config(B, n_gpus):
    return {'workload': 'EnvironmentModel.rollout_out rollout, BASELINE config #3 per GPU',
            'task': TASK, 'mode': 'training (per-row ref path of 3)', 'batch_per_gpu': B, 'vehicles': V,
            'obs_dim': D, 'horizon': H, 'global_batch': B * n_gpus,
            'parallelism': 'rows sharded over %d GPU(s), no data-path collective' % n_gpus,
            'l2': 'obs ping-pong working set %.0f MB < 126 MB L2 (step-to-step reuse is inherent to the '
                  'rollout); L2 flushed with a 512 MB memset between timed rollouts' % (3 * B * 576 / 1e6)}


def make_inputs(B, seed):
    from env_build_b200 import synthetic as syn
    from env_build_b200.dynamics_and_models import build_path_tables
    rng = np.random.default_rng(seed)
    paths = build_path_tables(TASK)[0]
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, TASK, V, paths, ref)
    tape = syn.make_actions(rng, H, B)
    return paths, obs, ref, tape


def mode_list():
    from env_build_b200 import synthetic as syn
    from env_build_b200.endtoend_env_utils import VEHICLE_MODE_LIST
    return syn.tiled_mode_list(VEHICLE_MODE_LIST[TASK], V)


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    obs, ref, tape, paths, modes = args
    from oracle import crossroad_oracle as orc
    m = orc.EnvironmentModel(TASK, N_FUT, mode='training', veh_mode_list=modes, path_list=paths)
    m.reset(obs, ref)
    acc = 0.0
    for t in range(tape.shape[0]):
        res = m.rollout_out(tape[t])
        acc += float(res[1].sum())
    return acc


class CpuArm(object):
    """Rows split over a process pool (one oracle model per core; NumPy is single threaded here,
    like the reference's TF pinned to 1 intra-/inter-op thread, dynamics_and_models.py:22-23)."""

    def __init__(self, rows_per_core=1024, seed=20210313):
        import multiprocessing as mp
        self.cores = len(os.sched_getaffinity(0))
        self.rows = rows_per_core * self.cores
        paths, obs, ref, tape = make_inputs(self.rows, seed)
        modes = mode_list()
        self.jobs = [(obs[i::self.cores].copy(), ref[i::self.cores].copy(), tape[:, i::self.cores].copy(), paths, modes)
                     for i in range(self.cores)]
        self.pool = mp.get_context('fork').Pool(self.cores)

    def step(self):
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, self.jobs)
        return time.perf_counter() - t0

    def one_thread_rate(self):
        """env-steps/s of ONE process on its share (the reference pins TF to one thread)."""
        t0 = time.perf_counter()
        _cpu_worker(self.jobs[0])
        return self.jobs[0][0].shape[0] * H / (time.perf_counter() - t0)

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self):
        return '%d rows x H=%d per step (%d rows per core), same synthetic distribution and config as the GPU arm' % (
            self.rows, H, self.rows // self.cores)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    arm = CpuArm()
    for _ in range(args.warmup):
        arm.step()
    times = [arm.step() for _ in range(args.steps)]
    arm.close()
    total = sum(times)
    value = arm.rows * H * args.steps / total
    cfg = workload_config(args.batch, args.gpus)
    cfg['sample'] = arm.describe()
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': cfg, 'gpu_launches': 0,
            'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': arm.cores, 'kind': 'port',
                             'sample': arm.describe(),
                             'note': 'NumPy fp32 restatement of the reference algorithm (oracle/); the TF2 '
                                     'reference cannot be installed offline'},
            'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'note': 'nvidia-smi unavailable'}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [s for t, s in self.samples if t0 <= t <= t1] or [s for _, s in self.samples[-3:]]
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic():
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(p):
        try:
            return json.load(open(p)).get('dram_bytes_per_launch')
        except Exception:
            return None
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from env_build_b200 import _lib
    from env_build_b200.dynamics_and_models import EnvironmentModel
    from env_build_b200.rollout import RolloutGraph

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if _lib.needs_build():
        if local == 0:
            _lib.build()
        if world > 1:
            dist.barrier()
    B, K, W = args.batch, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    paths, obs, ref, tape = make_inputs(B, 20210313 * 1000 + rank)
    model = EnvironmentModel(TASK, N_FUT, mode='training', veh_mode_list=mode_list())
    runner = RolloutGraph(model, B, V, H)
    runner.load(obs, ref, tape)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, n_warm, n_timed, after=None):
        for _ in range(n_warm):
            flush.zero_()
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_timed)]
        barrier()
        if after is None:
            for a, b in ev:
                flush.zero_()                   # L2 flush, outside the event pair
                a.record()
                fn()
                b.record()
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in ev)
        else:                                   # work on several streams: one event pair around all steps
            a, b = ev[0]
            a.record()
            for _ in range(n_timed):
                fn()
            after()                             # joins the side stream into the timing stream
            b.record()
            barrier()
            ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- value: inputs resident in HBM, graph replay ----
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    ms = timed(runner.run, W, K)
    launches = runner.launches_per_run * K
    value = world * B * H * K / (ms / 1e3)
    launch_us = 1e3 * ms / launches
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_ENV_STEP * B / (launch_us * 1e-6) / 1e9

    # ---- e2e: public API, pinned host buffers, copies inside the timed region ----
    h_obs = torch.from_numpy(obs).pin_memory()
    h_ref = torch.from_numpy(ref).pin_memory()
    h_tape = torch.from_numpy(tape).pin_memory()
    h_out5 = torch.empty((H, 5, B), dtype=torch.float32).pin_memory()
    h_final = torch.empty((B, D), dtype=torch.float32).pin_memory()

    # three streams: H2D of the next rollout's inputs, kernels, D2H of the previous rollout's results
    side, copy_s = torch.cuda.Stream(), torch.cuda.Stream()
    stage = [dict(obs=torch.empty((B, D), device=dev), ref=torch.empty((B,), dtype=torch.int32, device=dev),
                  tape=torch.empty((H, B, 2), device=dev), free=torch.cuda.Event(), ready=torch.cuda.Event())
             for _ in range(2)]
    counter = [0]

    def e2e_step():
        main = torch.cuda.current_stream()
        st = stage[counter[0] % 2]
        counter[0] += 1
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(st['free'])                     # the kernels that read this set are done
            st['obs'].copy_(h_obs, non_blocking=True)
            st['ref'].copy_(h_ref, non_blocking=True)
            st['tape'].copy_(h_tape, non_blocking=True)
            st['ready'].record(copy_s)
        main.wait_event(st['ready'])
        model.reset(st['obs'], st['ref'])
        outs = []
        for t in range(H):
            res = model.rollout_out(st['tape'][t])
            outs.append(model.last_out5)
        st['free'].record(main)
        out_all, final = torch.stack(outs), res[0]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            h_out5.copy_(out_all, non_blocking=True)
            h_final.copy_(final, non_blocking=True)
        out_all.record_stream(side)
        final.record_stream(side)

    ke = max(3, min(K, 20))
    def join_streams():
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.current_stream().wait_stream(copy_s)

    # the per-call API is host-driven (25 Python calls per rollout), so one descheduled host thread
    # shows up here: time the same ke steps three times and report the median, all three listed
    e2e_runs = sorted(timed(e2e_step, 3 if i == 0 else 1, ke, after=join_streams) for i in range(3))
    ms_e = e2e_runs[1]
    e2e_value = world * B * H * ke / (ms_e / 1e3)
    h2d = obs.nbytes + ref.nbytes + tape.nbytes
    d2h = h_out5.numel() * 4 + h_final.numel() * 4

    # ---- HBM-bound variant: batch too large for L2 (reported beside the headline, not as it) ----
    extra = {}
    if args.large_batch and world == 1:
        Bl = args.large_batch
        _, obs_l, ref_l, tape_l = make_inputs(Bl, 99)
        rl = RolloutGraph(model, Bl, V, H)
        rl.load(obs_l, ref_l, tape_l)
        ms_l = timed(rl.run, 3, 5)
        us_l = 1e3 * ms_l / (5 * H)
        extra = {'batch': Bl, 'value': Bl * H * 5 / (ms_l / 1e3), 'launch_us': us_l,
                 'achieved_gbs': BYTES_PER_ENV_STEP * Bl / (us_l * 1e-6) / 1e9,
                 'frac': BYTES_PER_ENV_STEP * Bl / (us_l * 1e-6) / 1e9 / peak,
                 'note': 'same kernel, batch whose ping-pong buffers (%.0f MB) exceed L2' % (2 * Bl * 576 / 1e6)}
        del rl

    # ---- optional MUFU sin/cos for the vehicles (ce2e_set_fast_trig): reported beside, never as `value`
    fast = None
    if world == 1:
        _lib.set_fast_trig(True)
        rf = RolloutGraph(model, B, V, H)
        rf.load(obs, ref, tape)
        ms_f = timed(rf.run, 3, K)
        _lib.set_fast_trig(False)
        fast = {'value': B * H * K / (ms_f / 1e3), 'launch_us': 1e3 * ms_f / (K * H),
                'frac': BYTES_PER_ENV_STEP * B / (1e3 * ms_f / (K * H) * 1e-6) / 1e9 / peak,
                'note': 'vehicle sin/cos from the special-function unit (abs err 2^-21.4); within the 1e-5 '
                        'tolerance but not the default build; NOT the headline'}
        del rf
    # ---- open-loop horizon-fused mode (ce2e_rollout_horizon): a DIFFERENT mode, reported apart
    fused = None
    if world == 1:
        rh = RolloutGraph(model, B, V, H, fused=True)
        rh.load(obs, ref, tape)
        ms_h = timed(rh.run, 3, K)
        fused = {'value': B * H * K / (ms_h / 1e3), 'unit': 'env-steps/s', 'ms_per_rollout': ms_h / K,
                 'note': 'one launch for all %d steps of an open-loop action tape, tile state resident on chip; '
                         'per-step bytes are only actions + the five outputs, so the per-step HBM roofline does '
                         'not apply; results bit-identical to the per-step launches; NOT the headline' % H}
        del rh
    # ---- N > 1: the batch lives on rank 0; NCCL scatters the row blocks and gathers the returns
    sharded = None
    if world > 1:
        from env_build_b200.parallel import ShardedRollout
        Bg = world * B
        sr = ShardedRollout(lambda b: RolloutGraph(model, b, V, H), Bg, D, H, dev)
        if rank == 0:
            _, g_obs, g_ref, g_tape = make_inputs(Bg, 4242)
            full = [torch.from_numpy(g_obs).to(dev), torch.from_numpy(g_ref).to(dev), torch.from_numpy(g_tape).to(dev)]
        else:
            full = [None, None, None]
        ret = {}

        def sharded_step():
            sr.scatter(*full)
            sr.run()
            ret['r'] = sr.gather_returns()

        ks = max(3, min(K, 10))
        ms_s = timed(sharded_step, 2, ks)
        sharded = {'value': Bg * H * ks / (ms_s / 1e3), 'unit': 'env-steps/s', 'ms_per_step': ms_s / ks,
                   'scatter_bytes_per_step': int(Bg * (D * 4 + 4 + H * 8)), 'gather_bytes_per_step': int(Bg * 20),
                   'note': 'global batch of %d rows resident on rank 0: NCCL scatter of observations, path indexes '
                           'and the action tape, %d-step rollout on every rank, NCCL gather of the per-row returns '
                           '(sum over steps of the five outputs)' % (Bg, H)}
        del sr, full
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None     # sampled over all GPU legs above
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        arm = CpuArm()
        arm.step()
        n, tot = 0, 0.0
        while tot < 10.0 and n < 20:
            tot += arm.step()
            n += 1
        one = arm.one_thread_rate()
        arm.close()
        cpu = {'value': arm.rows * H * n / tot, 'unit': 'env-steps/s', 'cores': arm.cores, 'kind': 'port',
               'value_1_thread': one,
               'sample': '%d steps of %s' % (n, arm.describe()),
               'note': 'NumPy fp32 restatement of the reference algorithm (oracle/), one process per core'}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(B, world),
                'clocks': clocks, 'gpu_launches': launches,
                'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': d2h, 'steps': ke, 'timing': 'one CUDA-event pair around all steps; H2D, kernels and D2H run on '
                        'three streams, all joined before the end event; median of three such runs',
                        'runs': [world * B * H * ke / (m / 1e3) for m in e2e_runs],
                        'path': 'EnvironmentModel.reset + %d x rollout_out; observations, path indexes and the action tape come from pinned host buffers, all per-step outputs and the final observations go back to pinned host buffers' % H},
                'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                             'frac': achieved / peak, 'traffic': ncu_traffic(), 'peak_source': peak_src,
                             'kernel': 'k_model_step<REW=1,NEXT=1>', 'launch_us': launch_us,
                             'bytes_per_launch': BYTES_PER_ENV_STEP * B,
                             'note': 'algorithmic bytes (8*D+32)*B per launch; at this batch the obs ping-pong '
                                     'fits L2, see large_batch for the HBM-bound rate'},
                'large_batch': extra or None, 'fast_trig_option': fast, 'horizon_fused_mode': fused,
                'sharded_from_rank0': sharded,
                'cpu_baseline': cpu}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints while the
    bench runs (NCCL's version banner, warnings) is diverted to stderr."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=65536, help='rows per GPU')
    ap.add_argument('--large-batch', type=int, default=524288)
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
