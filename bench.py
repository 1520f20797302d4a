#!/usr/bin/env python
"""bench.py -- expanded states/sec of the frontier-expansion path (BASELINE.json's metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One "step" = one complete beam search (root -> first state with >= goal points): every level
runs generate + dedup + score + top-k.  Workload at N=1 = BASELINE.json configs[2]:
speedrun, goal 15, `aggressive` heuristic, beam 30 M (`--beam` overrides), noise policy
`const`, ties by arrival order (`stable`).  The inputs are synthetic by construction: the
whole search grows from the all-zero root state and the rules' constant tables.

    value      expanded states / s, timed on the device (CUDA events) over K solves whose
               state lives in HBM throughout (spl_solver_* on a pre-built context)
    e2e        the same metric through the public API `State.newgame().solve(...)` with host
               inputs/outputs: root record H2D, per-level counters and the winning line D2H,
               path replay -- wall clock around the call
    roofline   dominant kernel (expand_kernel<PROBE>): algorithmic bytes / CUDA-event time
    cpu_baseline  the CPU oracle (C port of the reference algorithm, 1 thread) on a bounded
               sample of the same workload (same goal/heuristic/policy, beam 300 k)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

R_BYTES, K_BYTES = 24, 16  # record = key + aux, key
CPU_SAMPLE_BEAM = 300_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--goal', type=int, default=15)
    ap.add_argument('--heuristic', default='aggressive')
    ap.add_argument('--beam', type=int, default=30_000_000)
    ap.add_argument('--noise', default='const')
    ap.add_argument('--tie', default='stable')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw'

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 6] or [r for _, r in self.rows if len(r) >= 6]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm = sorted(int(float(r[0])) for r in rows)
        reasons = [n for i, n in enumerate(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], 2)
                   if any(r[i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': int(float(rows[0][1])), 'reasons': reasons, 'samples': len(rows)}


# ------------------------------------------------------------------ CPU arm (oracle port of the reference algorithm)
def cpu_solve(goal, heuristic, beam, tie, noise):
    import oracle
    t0 = time.perf_counter()
    s = oracle.Solver(goal, use_heuristic=True, heuristic_name=heuristic, beam_width=beam, policy=tie, noise=noise)
    infos = s.run()
    dt = time.perf_counter() - t0
    s.close()
    return sum(i['expanded'] for i in infos), dt


def run_reference(a):
    """`--impl reference`: the reference's algorithm on the host cores.  The reference is Python and
    cannot travel to the GPU box (only /root/repo does), so the arm times the oracle's C port of it
    -- single-threaded, like the reference (its search loop is sequential by construction)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    beam = min(a.beam, CPU_SAMPLE_BEAM)
    steps, warm = max(1, min(a.steps, 3)), min(a.warmup, 1)
    for _ in range(warm):
        cpu_solve(a.goal, a.heuristic, min(beam, 20_000), a.tie, a.noise)
    exp = tot = 0.0
    for _ in range(steps):
        e, dt = cpu_solve(a.goal, a.heuristic, beam, a.tie, a.noise)
        exp += e
        tot += dt
    v = exp / tot
    sample = f'goal {a.goal}, {a.heuristic}, beam {beam} (bounded sample of the beam-{a.beam} workload), {a.tie}/{a.noise}'
    print(json.dumps({
        'impl': 'reference', 'metric': 'expanded states/sec (gen+dedup+score+top-k)', 'value': v,
        'unit': 'expanded states/s', 'n_gpus': a.gpus, 'steps': steps, 'warmup': warm,
        'ms_per_step': tot / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u64 keys / f64 scores', 'data': 'synthetic',
        'config': {'workload': f'C3 speedrun goal {a.goal} -u -H {a.heuristic}, CPU-bounded sample beam {beam}'},
        'cpu_baseline': {'value': v, 'unit': 'expanded states/s', 'cores': 1, 'kind': 'port', 'sample': sample,
                         'host_cores': os.cpu_count()},
        'e2e': {'value': v, 'unit': 'expanded states/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ------------------------------------------------------------------ GPU arm
SHARDED_BEAM_PER_GPU = 12_500_000  # N > 1: beam = 12.5 M per GPU (N = 8 -> 100 M, BASELINE configs[3]'s width)


def run_b200(a):
    import torch
    import torch.distributed as dist

    import splendor_rl_gym_b200 as S
    from splendor_rl_gym_b200.sharded import Comm, CudaBackend, ShardedSolver

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        if a.beam == 30_000_000:  # default: keep the per-GPU queue fixed (weak scaling)
            a.beam = SHARDED_BEAM_PER_GPU * world
    # visited table sized for the whole search up front (about 85 visited states per beam slot at
    # goal 15) so the timed region never rehashes.  Three slots per 64-byte bucket; capped at
    # 2^30 buckets = 3.22 G slots = 68.7 GB (u32 slot ids), which also keeps the table inside the GPU's TLB
    # reach: beyond ~75 GB every random probe pays a page walk (profiles/README.md, r1c).
    slots = int(min(3 << 30, max(1 << 22, a.beam * 72 / 0.62 / world)))
    eng = S.Engine(local, table_slots=slots, max_table_bytes=int(150e9))
    k, aux = S.State.newgame().record()
    comm = Comm(eng.tdev)

    def solve_device():
        if world == 1:
            sol = eng.solver(k, aux, a.goal, True, a.heuristic, a.beam, a.tie, a.noise)
            infos = sol.run()
            sol.close()
        else:  # frontier sharded by key hash over the ranks; every level bit-identical to world == 1
            sol = ShardedSolver(CudaBackend(eng), comm, k, aux, a.goal, True, a.heuristic, a.beam, a.tie, a.noise, keep_links=False)
            infos = sol.run()
        return infos

    for _ in range(a.warmup):
        infos = solve_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record()
    all_infos = []
    for _ in range(a.steps):
        all_infos.append(solve_device())
    ev1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    tm = torch.tensor([ms], dtype=torch.float64, device='cuda')
    # per-level counters of the sharded solver are already global; count them once
    expanded = torch.tensor([float(sum(i['expanded'] for inf in all_infos for i in inf))], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        lt = torch.tensor([float(launches)], dtype=torch.float64, device='cuda')
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms = float(tm.item())
    value = float(expanded.item()) / (ms * 1e-3)
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    # ---- end to end through the public API (host inputs / outputs inside the timed region)
    h0, d0 = eng.transfer_bytes()
    st = []
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    te0 = time.perf_counter()
    for _ in range(a.steps):
        path = S.State.newgame().solve(goal_pts=a.goal, use_heuristic=True, heuristic_name=a.heuristic,
                                       beam_width=a.beam, verbose=False, tie_policy=a.tie, noise=a.noise,
                                       engine=eng, stats=st)
    torch.cuda.synchronize()
    te = time.perf_counter() - te0
    h1, d1 = eng.transfer_bytes()
    e2e_t = torch.tensor([te], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_exp = sum(i['expanded'] for i in st)
    # bytes that crossed PCIe per solve: root record + per-launch scalars up; counters, select state,
    # parent links and the replayed successor lists of the winning line (<= 190 * 40 B per move) down
    moves = len(path) - 1
    e2e = {'value': float(e2e_exp) / float(e2e_t.item()), 'unit': 'expanded states/s',
           'h2d_bytes_per_step': (h1 - h0) // a.steps + moves * 24,
           'd2h_bytes_per_step': (d1 - d0) // a.steps + moves * 190 * 40,
           'moves': moves, 'final': repr(path[-1]), 'final_pts': path[-1].pts,
           'seconds_per_solve': float(e2e_t.item()) / a.steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from the last timed solve's per-level CUDA-event times
    infos = all_infos[-1]
    lv = [i for i in infos if i['expanded'] and i['generated']]
    if world > 1:
        line = {
            'metric': 'expanded states/sec (gen+dedup+score+top-k)', 'value': value, 'unit': 'expanded states/s',
            'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64 keys / f64 scores',
            'data': 'synthetic',
            'config': {'workload': f'C3 (BASELINE configs[2]): speedrun goal {a.goal} -u -H {a.heuristic}, beam {a.beam} '
                                   f'(= {a.beam // world} per GPU), noise={a.noise}, ties={a.tie}; frontier sharded by key hash, '
                                   f'NCCL all-to-all routing; one step = one full solve',
                       'l2': 'working set (visited table %.1f GB per GPU) >> 126 MB L2; no flush needed' % (slots // 3 * 64 / 1e9),
                       'beam': a.beam, 'goal': a.goal, 'heuristic': a.heuristic, 'parallelism': f'hash-sharded x{world}'},
            'time_to_solve_goal15_s': ms / a.steps * 1e-3, 'levels': len(infos),
            'expanded_per_step': sum(i['expanded'] for i in lv), 'generated_per_step': sum(i['generated'] for i in lv),
            'unique_per_step': sum(i['unique'] for i in lv), 'visited': infos[-2]['visited'] if len(infos) > 1 else None,
            'roofline': None, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks,
            'note': 'roofline / cpu_baseline are reported by the N=1 line (same kernels); this line is the sharded driver',
        }
        print(json.dumps(line))
        dist.destroy_process_group()
        return
    ms_stage = {s: sum(i['ms_' + s] for i in lv) for s in ('count', 'expand', 'resolve', 'select', 'sort')}
    n = sum(i['expanded'] for i in lv)
    gen = sum(i['generated'] for i in lv)
    uniq = sum(i['unique'] for i in lv)
    kept = sum(i['kept'] for i in lv)
    # expand_kernel<PROBE>: read each parent (R), one visited-table key read per candidate (K),
    # one table insert per new unique (R)                       [SURVEY.md 8(d) terms R + K*b + R*u]
    expand_bytes = R_BYTES * n + K_BYTES * gen + R_BYTES * uniq
    chunk = 4 << 20  # spl_config.chunk_parents default: one expand launch per 4 Mi parents of a level
    n_exp_launches = max(1, sum(-(-i['frontier'] // chunk) for i in lv))
    peaks = {}
    try:
        peaks = json.load(open(ROOT / 'MEASURED_PEAKS.json'))
    except OSError:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    achieved = expand_bytes / (ms_stage['expand'] * 1e-3) / 1e9
    traffic = None
    prof = ROOT / 'profiles' / 'expand_kernel_traffic.json'
    if prof.exists():
        try:
            traffic = json.load(open(prof)).get('dram_bytes_per_launch')
        except (OSError, ValueError):
            traffic = None
    # whole pipeline against the SURVEY formula: R + K*b + (2R+16)*u + 2R*k per expanded state
    pipe_bytes = R_BYTES * n + K_BYTES * gen + (2 * R_BYTES + 16) * uniq + 2 * R_BYTES * kept
    pipe_ms = sum(ms_stage.values())
    line = {
        'metric': 'expanded states/sec (gen+dedup+score+top-k)', 'value': value, 'unit': 'expanded states/s',
        'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64 keys / f64 scores',
        'data': 'synthetic',
        'config': {'workload': f'C3 (BASELINE configs[2]): speedrun goal {a.goal} -u -H {a.heuristic}, beam {a.beam}, '
                               f'noise={a.noise}, ties={a.tie}; one step = one full solve from the root state',
                   'l2': 'working set (visited table %.1f GB) >> 126 MB L2; no flush needed' % (slots // 3 * 64 / 1e9),
                   'beam': a.beam, 'goal': a.goal, 'heuristic': a.heuristic, 'parallelism': 'single GPU, fused expand+probe kernels'},
        'time_to_solve_goal15_s': ms / a.steps * 1e-3,
        'generated_per_s': float(gen) * a.steps / (ms * 1e-3) if world == 1 else None,
        'levels': len(infos), 'expanded_per_step': n, 'generated_per_step': gen, 'unique_per_step': uniq,
        'visited': infos[-1]['visited'],
        'stage_ms_per_step': {k_: round(v, 3) for k_, v in ms_stage.items()},
        'roofline': {'bound': 'hbm', 'kernel': 'expand_kernel<MODE_PROBE>', 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'traffic_note': 'ncu dram bytes of one saturated 4 Mi-parent launch (profiles/expand_kernel_traffic.json); '
                                     'its algorithmic bytes are 2.6e9',
                     'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s (B200_PROFILING.md)',
                     'algorithmic_bytes_per_launch': expand_bytes / n_exp_launches,
                     'avg_launch_ms': ms_stage['expand'] / n_exp_launches, 'launches': n_exp_launches,
                     'pipeline_frac': (pipe_bytes / (pipe_ms * 1e-3) / 1e9) / peak},
        'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks,
    }
    if not a.no_cpu_baseline:
        cexp, cdt = cpu_solve(a.goal, a.heuristic, min(a.beam, CPU_SAMPLE_BEAM), a.tie, a.noise)
        line['cpu_baseline'] = {'value': cexp / cdt, 'unit': 'expanded states/s', 'cores': 1, 'kind': 'port',
                                'host_cores': os.cpu_count(),
                                'sample': f'oracle C port, goal {a.goal} {a.heuristic} beam {min(a.beam, CPU_SAMPLE_BEAM)} '
                                          f'({cexp} expanded in {cdt:.1f} s); the Python reference itself measured '
                                          f'6.4 k expanded/s on this workload (BASELINE.md 2)'}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)
