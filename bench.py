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
  e2e           the same metric from pinned HOST buffers through the public RolloutGraph API (load ->
                run -> results to the host: H2D of observations, path indexes and the action tape, D2H of
                every step's 5 outputs and the final observations inside the timed region); sub-records:
                per_step_api (EnvironmentModel.reset + 25 x rollout_out), shield_outputs (D2H of the
                per-step veh2veh4real only), link_ceiling (only the copies)
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
    return {'workload': 'EnvironmentModel.rollout_out rollout, BASELINE config #3 per GPU (weak scaling; the N=8 line '
                        'also carries config #5 at its stated size under "config5")',
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

    # rollouts per timed e2e run (its own count, reported as e2e.steps): the run starts with an empty pipeline and
    # ends drained, so fewer than 20 rollouts would mostly measure the fill
    ke = max(20, min(K, 40))
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

    # the same rollout through the other public entry, RolloutGraph (static device buffers, the 25 launches
    # replayed as one CUDA graph): load() from the pinned host buffers, run(), results back to pinned host
    # buffers; three buffer sets so that H2D, kernels and D2H of consecutive rollouts overlap (2 / 3 / 4 sets measured:
    # 1.085-1.099e9 / 1.104-1.106e9 / 1.087-1.110e9 on one box)
    NSET = max(2, int(os.environ.get('CE2E_E2E_SETS', '3')))          # buffer sets of the e2e pipeline
    graphs = [RolloutGraph(model, B, V, H) for _ in range(NSET)]
    for g_ in graphs:
        g_.load(obs, ref, tape)
        g_.run()
    torch.cuda.synchronize()
    gev = [dict(free=torch.cuda.Event(), ready=torch.cuda.Event(), done=torch.cuda.Event(), out=torch.cuda.Event())
           for _ in range(NSET)]
    gcount = [0]

    def e2e_graph_step():
        main = torch.cuda.current_stream()
        k = gcount[0] % NSET
        gcount[0] += 1
        g_, ev_ = graphs[k], gev[k]
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_['free'])                    # the rollout that last read these inputs is done
            g_.load(h_obs, h_ref, h_tape)
            ev_['ready'].record(copy_s)
        main.wait_event(ev_['ready'])
        main.wait_event(ev_['out'])                           # its previous results have left the device
        g_.run()
        ev_['free'].record(main)
        ev_['done'].record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev_['done'])
            h_out5.copy_(g_.out5, non_blocking=True)
            h_final.copy_(g_.final_obs, non_blocking=True)
            ev_['out'].record(side)

    for ev_ in gev:
        ev_['free'].record(torch.cuda.current_stream())
        ev_['out'].record(torch.cuda.current_stream())
    gr_runs = sorted(timed(e2e_graph_step, 3 if i == 0 else 1, ke, after=join_streams) for i in range(3))
    e2e_graph = {'value': world * B * H * ke / (gr_runs[1] / 1e3), 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d,
                 'd2h_bytes_per_step': d2h, 'runs': [world * B * H * ke / (m / 1e3) for m in gr_runs],
                 'path': 'RolloutGraph.load(pinned host obs / path indexes / action tape) + run() (one CUDA graph of %d '
                         'ce2e_rollout_step launches) + D2H of all per-step outputs and the final observations' % H}
    del graphs

    # the same path when the caller consumes what the reference's shield rollouts consume: the per-step
    # veh2veh4real vector only (hier_decision.py:97, multi_ego.py:197); same H2D, 1/10 of the D2H
    h_v2v = torch.empty((H, B), dtype=torch.float32).pin_memory()

    def e2e_shield_step():
        main = torch.cuda.current_stream()
        st = stage[counter[0] % 2]
        counter[0] += 1
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(st['free'])
            st['obs'].copy_(h_obs, non_blocking=True)
            st['ref'].copy_(h_ref, non_blocking=True)
            st['tape'].copy_(h_tape, non_blocking=True)
            st['ready'].record(copy_s)
        main.wait_event(st['ready'])
        model.reset(st['obs'], st['ref'])
        v2v = st['v2v']
        for t in range(H):
            v2v[t] = model.rollout_out(st['tape'][t])[4]
        st['free'].record(main)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            h_v2v.copy_(v2v, non_blocking=True)

    for st_ in stage:
        st_['v2v'] = torch.empty((H, B), device=dev)
    sh_runs = sorted(timed(e2e_shield_step, 2 if i == 0 else 1, ke, after=join_streams) for i in range(3))
    e2e_shield = {'value': world * B * H * ke / (sh_runs[1] / 1e3), 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d,
                  'd2h_bytes_per_step': h_v2v.numel() * 4,
                  'note': 'same public-API path, D2H of the per-step veh2veh4real vector only (what the reference\'s '
                          'safety-shield callers read, hier_decision.py:97); NOT the headline e2e'}

    # what the box's host<->device links give for exactly these transfers, nothing else running: every
    # rank moves the e2e leg's H2D and D2H bytes concurrently from / to its pinned buffers
    def bare_copies():
        st = stage[0]
        with torch.cuda.stream(copy_s):
            st['obs'].copy_(h_obs, non_blocking=True)
            st['ref'].copy_(h_ref, non_blocking=True)
            st['tape'].copy_(h_tape, non_blocking=True)
        with torch.cuda.stream(side):
            h_out5.copy_(d_out5, non_blocking=True)
            h_final.copy_(st['obs'], non_blocking=True)

    d_out5 = torch.empty((H, 5, B), device=dev)
    ms_c = timed(bare_copies, 2, ke, after=join_streams)
    link = {'value': world * B * H * ke / (ms_c / 1e3), 'unit': 'env-steps/s',
            'h2d_gbs_per_gpu': h2d * ke / (ms_c / 1e3) / 1e9, 'd2h_gbs_per_gpu': d2h * ke / (ms_c / 1e3) / 1e9,
            'aggregate_gbs': world * (h2d + d2h) * ke / (ms_c / 1e3) / 1e9,
            'note': 'ceiling of the e2e leg on this box: only its pinned-memory H2D + D2H copies, all %d rank(s) at '
                    'once, no kernels' % world}

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
    # ---- the reference's own vehicle counts (EU:21-23, EU:40-42): one launch per step at V = 8 / 9 / 5
    native = None
    if world == 1:
        from env_build_b200 import synthetic as syn
        from env_build_b200.endtoend_env_utils import VEHICLE_MODE_LIST, VEH_NUM
        native = {}
        for task_n in ('left', 'straight', 'right'):
            Vn = VEH_NUM[task_n]
            Dn = 9 + 4 * Vn
            rng_n = np.random.default_rng(5)
            mn = EnvironmentModel(task_n, 0, mode='training', veh_mode_list=VEHICLE_MODE_LIST[task_n])
            ref_n = syn.make_ref_indexes(rng_n, B)
            rn = RolloutGraph(mn, B, Vn, H)
            rn.load(syn.make_obs(rng_n, B, task_n, Vn, mn.ref_path.path_list, ref_n), ref_n, syn.make_actions(rng_n, H, B))
            ms_n = timed(rn.run, 3, K)
            us_n = 1e3 * ms_n / (K * H)
            native[task_n] = {'vehicles': Vn, 'obs_dim': Dn, 'value': B * H * K / (ms_n / 1e3), 'launch_us': us_n,
                              'bytes_per_env_step': 8 * Dn + 32,
                              'frac': (8 * Dn + 32) * B / (us_n * 1e-6) / 1e9 / peak}
            del rn, mn
        native['note'] = 'same kernel and batch at the task\'s native vehicle count; %d rows x 8D+32 bytes is %s of ' \
                         'HBM time, so these launches are bound by per-launch latency, not bandwidth' % (B, '3-4 us')
    # ---- f-1: the batched SUMO-free CrossroadEnd2end, one step = fused model step + done kernel + on-device
    # auto-reset, replayed as one CUDA graph (no host synchronisation)
    env_rec = None
    if world == 1:
        from env_build_b200.endtoend import CrossroadEnd2end
        env = CrossroadEnd2end(TASK, num_envs=B, veh_num=V, auto_reset=True, use_graph=True, reward_info=False)
        env.seed(1)
        env.reset()
        a_env = torch.zeros((B, 2), device=dev)
        for _ in range(4):
            env.step(a_env)
        n_env = 100
        ms_env = timed(lambda: [env.step(a_env) for _ in range(n_env)], 1, 3)
        env_rec = {'us_per_step': 1e3 * ms_env / (3 * n_env), 'value': B * 3 * n_env / (ms_env / 1e3), 'unit': 'env-steps/s',
                   'note': 'CrossroadEnd2end(num_envs=%d, veh_num=%d, auto_reset=True, use_graph=True).step(): two '
                           'kernels per step (k_model_step_pair; k_env_done = done logic + restart of finished rows) in one graph replay, Python '
                           'call included' % (B, V)}
        del env
    # ---- N > 1: the batch lives on rank 0; NCCL scatters the row blocks and gathers the returns
    sharded, config5 = None, None
    if world > 1:
        from env_build_b200.parallel import ShardedRollout

        def sharded_leg(Bper, check):
            """Global batch of world * Bper rows resident on rank 0 in scatter-ready layout; per rollout: the
            exchange (blocks of it -> the ranks' static buffers), H-step rollout on every rank, NCCL gather of the
            per-row returns.  Three buffer sets per rank: the next rollouts' inputs travel on a side stream under
            the current rollout's kernels.  The timed region is self-contained (pipeline fill included)."""
            Bg = world * Bper
            # transport of the staged batch: copy-engine pulls over NVLink peer mappings when the box offers
            # them (every rank must agree), else NCCL's scatter
            sr, exchange = None, os.environ.get('CE2E_EXCHANGE', 'peer')
            S = max(2, int(os.environ.get('CE2E_EXCHANGE_SLOTS', '3')))    # buffer sets per rank
            if exchange == 'peer':
                try:
                    sr = ShardedRollout(lambda b: RolloutGraph(model, b, V, H), Bg, D, H, dev, slots=S, exchange='peer')
                except Exception as e:                                   # noqa: BLE001
                    print('[bench] peer exchange unavailable on rank %d: %s' % (rank, str(e)[:200]), file=sys.stderr)
                agree = torch.tensor([int(sr is not None)], device=dev)
                dist.all_reduce(agree, op=dist.ReduceOp.MIN)
                if not int(agree):
                    sr, exchange = None, 'nccl'
            if sr is None:
                sr = ShardedRollout(lambda b: RolloutGraph(model, b, V, H), Bg, D, H, dev, slots=S)
            staged = None
            if rank == 0:
                _, g_obs, g_ref, g_tape = make_inputs(Bg, 4242)
                staged = sr.stage(g_obs, g_ref, g_tape)
            for r_ in sr.runners:
                r_.run()                                  # capture the graphs before any timing
            comm, gath = torch.cuda.Stream(), torch.cuda.Stream()
            ready = [torch.cuda.Event() for _ in range(S)]
            done = [torch.cuda.Event() for _ in range(S)]
            gathered = [torch.cuda.Event() for _ in range(S)]
            ret = {}
            state = {'i': 0}

            def issue_scatter(slot):
                with torch.cuda.stream(comm):
                    comm.wait_event(done[slot])           # the rollout that last used this buffer set is finished
                    sr.scatter_staged(staged, slot=slot)
                    ready[slot].record(comm)

            def region(n):
                # n rollouts, self-contained (n exchanges, n rollouts, n gathers; nothing in flight before or after):
                # rollout i computes on buffer set i % S while the inputs of rollouts i+1 .. i+S-1 travel -- with three
                # sets the exchange stream always has the next transfer queued behind the current one
                main = torch.cuda.current_stream()
                for e_ in done + gathered:
                    e_.record(main)
                gath.wait_stream(main)
                for k_ in range(min(S - 1, n)):
                    issue_scatter(k_)
                for i in range(n):
                    slot = i % S
                    if i + S - 1 < n:
                        issue_scatter((i + S - 1) % S)
                    main.wait_event(ready[slot])
                    main.wait_event(gathered[slot])       # the returns of the rollout that last used this set are out
                    sr.run(slot=slot)
                    done[slot].record(main)
                    with torch.cuda.stream(gath):         # the returns' reduction + gather leave the kernel stream
                        gath.wait_event(done[slot])
                        ret['r'] = sr.gather_returns(slot=slot)
                        gathered[slot].record(gath)

            def join():
                torch.cuda.current_stream().wait_stream(comm)
                torch.cuda.current_stream().wait_stream(gath)

            ks = max(6, min(2 * K, 20))
            ms_s = timed(lambda: region(ks), 1, 1, after=join)
            scatter_bytes = int(world * sr.runner.inbox.numel() * 4)
            out = {'value': Bg * H * ks / (ms_s / 1e3), 'unit': 'env-steps/s', 'ms_per_step': ms_s / ks,
                   'exchange': {'peer': 'peer pull: every rank copies its block out of rank 0\'s NVLink-mapped staged '
                                        'batch with the copy engines (no send/recv kernels)',
                                'nccl': 'NCCL scatter'}[exchange],
                   'global_batch': Bg, 'scatter_bytes_per_step': scatter_bytes, 'gather_bytes_per_step': int(Bg * 20),
                   'root_egress_gbs': scatter_bytes * (world - 1) / world / (ms_s / ks / 1e3) / 1e9,
                   'nvlink_per_direction_gbs': {'nominal': 900, 'measured_peer_copy': 770, 'measured_pull_7_peers': 835},
                   'buffer_sets': S, 'rollouts_timed': ks,
                   'timing': 'one event pair around a self-contained run of rollouts_timed rollouts (first exchange not '
                             'overlapped), max over ranks'}
            if check:
                # driver-side NCCL correctness: the gathered returns against a local rollout of the same rows
                torch.cuda.synchronize()
                verdict = 'n/a'
                if rank == 0:
                    got = ret['r']
                    ok = True
                    one = RolloutGraph(model, Bper, V, H)
                    _, g_obs, g_ref, g_tape = make_inputs(Bg, 4242)
                    for r_ in range(world):
                        lo, hi = r_ * Bper, (r_ + 1) * Bper
                        one.load(g_obs[lo:hi], g_ref[lo:hi], g_tape[:, lo:hi])
                        one.run()
                        ok = ok and bool(torch.equal(one.out5.sum(0).t().contiguous(), got[lo:hi]))
                    verdict = 'bit-identical' if ok else 'MISMATCH'
                    del one
                out['sharded_check'] = verdict
            del sr, staged
            return out

        sharded = sharded_leg(B, True)
        sharded['vs_exchange_free'] = sharded['value'] / value
        sharded['note'] = ('global batch resident on rank 0 in scatter-ready layout (padded observation rows, action tape, '
                           'path indexes per rank block); per rollout: the exchange named in "exchange" into the ranks\' '
                           'static buffers, %d-step rollout on every rank, NCCL gather of the per-row returns (sum over '
                           'steps of the five outputs); three buffer sets per rank: the exchanges of rollouts i+1 and i+2 overlap the '
                           'kernels of rollout i' % H)
        if world == 8 or os.environ.get('CE2E_BENCH_CONFIG5') == '1':     # (the switch: exercise this block on fewer GPUs)
            # ---- BASELINE config #5 at its stated size: B = 1 048 576 over 8 GPUs (131 072 rows per GPU)
            B5 = 131072
            _, obs5, ref5, tape5 = make_inputs(B5, 20210315 * 1000 + rank)
            r5 = RolloutGraph(model, B5, V, H)
            r5.load(obs5, ref5, tape5)
            ms_5 = timed(r5.run, 3, max(3, min(K, 10)))
            k5 = max(3, min(K, 10))
            us_5 = 1e3 * ms_5 / (k5 * H)
            del r5
            config5 = {'workload': 'BASELINE config #5: batch=1048576 full model rollout horizon=25 on 8xB200',
                       'global_batch': world * B5, 'batch_per_gpu': B5, 'value': world * B5 * H * k5 / (ms_5 / 1e3),
                       'unit': 'env-steps/s', 'launch_us': us_5,
                       'frac_per_gpu': BYTES_PER_ENV_STEP * B5 / (us_5 * 1e-6) / 1e9 / peak,
                       'with_exchange': sharded_leg(B5, False)}
            config5['with_exchange']['vs_exchange_free'] = config5['with_exchange']['value'] / config5['value']
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
                'e2e': {'value': e2e_graph['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': d2h, 'steps': ke,
                        'timing': 'one CUDA-event pair around all steps; H2D, kernels and D2H run on three streams '
                                  '(three buffer sets), all joined before the end event; median of three such runs. The region '
                                  'starts with an empty pipeline and ends drained: the first H2D + rollout (1.4 ms) is not '
                                  'overlapped, after that a rollout costs its D2H (1.25 ms alone, ~1.35 ms under the '
                                  'concurrent H2D)',
                        'runs': e2e_graph['runs'], 'path': e2e_graph['path'],
                        'per_step_api': {'value': e2e_value, 'unit': 'env-steps/s',
                                         'runs': [world * B * H * ke / (m / 1e3) for m in e2e_runs],
                                         'path': 'EnvironmentModel.reset + %d x rollout_out (one Python call and three '
                                                 'fresh output tensors per step, like the reference\'s eager TF); same '
                                                 'pinned host buffers, same bytes; host-bound and sensitive to host '
                                                 'jitter, hence not the headline' % H},
                        'shield_outputs': e2e_shield, 'link_ceiling': link,
                        'vs_link_ceiling': e2e_graph['value'] / link['value']},
                'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                             'frac': achieved / peak, 'traffic': ncu_traffic(), 'peak_source': peak_src,
                             'kernel': 'k_model_step_pair (TMA tensor-map staging; the fused rollout_out step)',
                             'launch_us': launch_us,
                             'bytes_per_launch': BYTES_PER_ENV_STEP * B,
                             'note': 'algorithmic bytes (8*D+32)*B per launch; at this batch the obs ping-pong '
                                     'fits L2, see large_batch for the HBM-bound rate'},
                'large_batch': extra or None, 'native_v': native, 'env_step': env_rec, 'fast_trig_option': fast,
                'horizon_fused_mode': fused, 'sharded_from_rank0': sharded, 'config5': config5,
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
