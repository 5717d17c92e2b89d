#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: gates/s (and achieved HBM GB/s) of a depth-20 layered random
1q/2q circuit on a 30-qubit complex128 state per GPU (workload W-B of SURVEY.md 8d, config "Random circuit 30
qubits complex128 on 1xB200 with gate fusion"); with --gpus N the state has 30+log2(N) qubits, sharded on its
top qubits (weak scaling, BASELINE.json configs[4] shape).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--qubits Q] [--depth D] [--impl reference]

One step = one pass of the whole circuit over the state. Prints ONE JSON line (rank 0).
  value        device-timed (CUDA events on the launching stream, max over ranks), state resident in HBM
  e2e          the same circuit through the public API (Circuit.run on a State built from a pinned HOST buffer,
               result copied back to a pinned HOST buffer): host<->device copies inside the timed region
  roofline     dominant kernel = sweep_kernel (one launch = one read + one write of the state = 32 B/amplitude)
  cpu_baseline the C/OpenMP restatement of the reference's tensormul on the host cores, bounded sample
--impl reference times the reference's own algorithm (np.einsum with the reference's subscripts, one thread --
that is all numpy's einsum uses) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--qubits', type=int, default=None, help='qubits per GPU (default 30)')
    ap.add_argument('--depth', type=int, default=20)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--tile-bits', type=int, default=None)
    ap.add_argument('--low-bits', type=int, default=None)
    ap.add_argument('--max-cost', type=float, default=None)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=15.0, help='target wall time of the CPU baseline sample')
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.lines.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            inside = t0 - 0.1 <= ts <= t1 + 0.3
            try:
                if inside:
                    sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            if inside:
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic_bytes(nbits: int):
    """dram read+write bytes per sweep launch from the committed ncu --set full capture (profiles/), if any."""
    path = os.path.join(ROOT, 'profiles', 'sweep_traffic.json')
    try:
        with open(path) as f:
            rec = json.load(f)
        if int(rec.get('nbits', -1)) == nbits:
            return float(rec['dram_bytes_per_launch'])
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------------------
# CPU arms (oracle; rank 0 only)
# ---------------------------------------------------------------------------------------------------------

def cpu_baseline_c_port(nq: int, depth: int, seed: int, target_seconds: float) -> dict:
    """C/OpenMP restatement of the reference's tensormul on all host threads, bounded sample of the circuit."""
    import numpy as np
    from oracle import c_oracle
    from oracle import qf_oracle as O
    from quantumflow_b200 import workloads
    specs = workloads.wb_gate_list(nq, depth, seed)
    threads = c_oracle.max_threads()
    state = np.zeros(1 << nq, dtype=np.complex128)
    state[0] = 1.0
    done = 0
    t0 = time.perf_counter()
    for name, params, qubits in specs:
        c_oracle.apply_dense(state, O.gate_matrix(name, params), [nq - 1 - q for q in qubits])
        done += 1
        if time.perf_counter() - t0 > target_seconds:
            break
    dt = time.perf_counter() - t0
    return {'value': done / dt, 'unit': 'gates/s', 'cores': threads, 'kind': 'port',
            'sample': 'first {} of {} gates of the same {}-qubit circuit, C/OpenMP restatement of '
                      'numpybk.tensormul (oracle/qf_oracle_c.c), {:.1f} s'.format(done, len(specs), nq, dt)}


def reference_einsum_step(nq: int, specs, start: int, ngates: int, state):
    """`ngates` gates of the circuit with the reference's own np.einsum call (numpybk.py:159-214)."""
    from oracle import qf_oracle as O
    t0 = time.perf_counter()
    for i in range(ngates):
        name, params, qubits = specs[(start + i) % len(specs)]
        state = O.tensormul(O.as_tensor(O.gate_matrix(name, params)), state, list(qubits))
    return time.perf_counter() - t0, state


def run_reference_arm(args):
    """bench.py --impl reference: the reference's algorithm on the host (single-threaded np.einsum)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import numpy as np
    from quantumflow_b200 import workloads
    world = max(1, args.gpus)
    p = world.bit_length() - 1
    nq = (args.qubits or 30) + p
    specs = workloads.wb_gate_list(nq, args.depth, args.seed)
    # bounded sample: the circuit's gates are taken in order, `per_step` gates per step
    probe_n = min(nq, 24)
    probe = np.zeros([2] * probe_n, dtype=np.complex128)
    probe[(0,) * probe_n] = 1
    dt_probe, _ = reference_einsum_step(probe_n, workloads.wb_gate_list(probe_n, 1, 0), 0, 4, probe)
    est_gate = dt_probe / 4 * (2 ** (nq - probe_n))
    budget = 150.0 / max(1, args.steps + args.warmup)
    per_step = max(1, int(budget / max(est_gate, 1e-6)))
    per_step = min(per_step, len(specs))
    try:
        state = np.zeros([2] * nq, dtype=np.complex128)
        state[(0,) * nq] = 1
    except MemoryError:
        print(json.dumps({'impl': 'reference', 'unavailable': 'host cannot hold a {}-qubit state'.format(nq)}))
        return
    pos = 0
    for _ in range(args.warmup):
        _, state = reference_einsum_step(nq, specs, pos, per_step, state)
        pos += per_step
    total = 0.0
    for _ in range(args.steps):
        dt, state = reference_einsum_step(nq, specs, pos, per_step, state)
        pos += per_step
        total += dt
    gates_per_s = per_step * args.steps / total
    value = gates_per_s * world          # shard-gates/s, see main arm
    sample = ('{} consecutive gates of the {}-qubit circuit per step (np.einsum with the reference subscripts, '
              'single thread), {} steps'.format(per_step, nq, args.steps))
    line = {
        'impl': 'reference', 'metric': 'gates/s', 'value': value, 'unit': 'gates/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex128',
        'data': 'synthetic',
        'config': workload_config(nq, args.depth, args.seed, world, len(specs)),
        'cpu_baseline': {'value': value, 'unit': 'gates/s', 'cores': 1, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'gates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(nq, depth, seed, world, ngates):
    return {'workload': 'W-B depth-{} layered random 1q/2q circuit ({{H,X,T,RX,RY,RZ}} per qubit + CNOT/CZ on a '
                        'random perfect matching per layer), {} qubits, complex128, seed {}, {} gates'.format(
                            depth, nq, seed, ngates),
            'qubits': nq, 'depth': depth, 'gates': ngates, 'seed': seed,
            'sharding': 'none' if world == 1 else 'top {} qubits over {} GPUs'.format(world.bit_length() - 1, world),
            'unit_of_work': 'one gate applied to one 2^(qubits per GPU)-amplitude shard '
                            '(value = circuit gates/s x n_gpus; identical to plain gates/s at n_gpus=1)',
            'l2': 'state is {} MiB per GPU, far larger than the 126 MB L2 (no flush needed)'.format(
                (16 << (nq - (world.bit_length() - 1))) >> 20)}


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------

def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and args.gpus != world:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    import quantumflow_b200 as qf
    from quantumflow_b200 import engine, planner, workloads

    p = world.bit_length() - 1
    assert (1 << p) == world, 'number of GPUs must be a power of two'
    nlocal = args.qubits or 30
    nq = nlocal + p
    specs = workloads.wb_gate_list(nq, args.depth, args.seed)
    ngates = len(specs)

    t_plan0 = time.perf_counter()
    circ = workloads.circuit_from_specs(qf, specs)
    if world == 1:
        bitops = [(g.matrix(), [nq - 1 - q for q in g.qubits]) for g in circ.elements]
        segments = planner.build_segments(nq, bitops, tile_bits=args.tile_bits, low_bits=args.low_bits,
                                          max_cost=args.max_cost)
        runner = None
    else:
        from quantumflow_b200 import sharded
        runner = sharded.ShardedCircuit(circ, nq, world, rank, tile_bits=args.tile_bits, low_bits=args.low_bits,
                                        max_cost=args.max_cost)
        segments = runner.local_segments()
    plan_seconds = time.perf_counter() - t_plan0
    stats = planner.plan_stats(segments)
    # fingerprint of the executed plan (all sweep records): lets a reader check which plan a number belongs to
    import hashlib
    stats['plan_sha256'] = hashlib.sha256(b''.join(s.blob for s in segments if s.blob)).hexdigest()[:16]

    dev = torch.device('cuda', local_rank)
    state = torch.zeros(1 << nlocal, dtype=torch.complex128, device=dev)
    if rank == 0:
        state[0] = 1.0

    def step():
        nonlocal state
        if runner is None:
            qf.Circuit._execute(segments, state)
        else:
            state = runner.execute(state)     # in place (the remaps exchange blocks through staging chunks)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    if runner is not None:
        runner.reset_comm_counters()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = engine.launch_count()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t1 = time.time()
    launches = engine.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_step = ms_total / args.steps
    gates_per_s = ngates / (ms_step * 1e-3)
    value = gates_per_s * world

    # unitarity check after all steps: the state must still be normalised (work was really done, correctly)
    n2 = engine.norm2(state)
    if world > 1:
        dist.all_reduce(n2, op=dist.ReduceOp.SUM)
    norm_err = abs(float(n2) - 1.0)

    # roofline of the dominant kernel: sweep launches only (remap / fallback kernels are excluded from the count)
    nsweeps = stats['sweeps']
    comm_ms = runner.comm_ms_per_step() if runner is not None else 0.0
    sweep_ms = (ms_step - comm_ms) / max(1, nsweeps)
    algo_bytes = 32.0 * (1 << nlocal)
    peak, peak_src = measured_peak_gbs()
    achieved = algo_bytes / (sweep_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': ncu_traffic_bytes(nlocal), 'kernel': 'sweep_kernel<{}>'.format(
                    args.tile_bits or planner.DEFAULT_TILE_BITS),
                'algorithmic_bytes_per_launch': algo_bytes, 'launches_per_step': nsweeps,
                'avg_launch_ms': sweep_ms, 'peak_source': peak_src}

    # end-to-end through the public API with host buffers (N=1: State from pinned host memory -> Circuit.run ->
    # result back in pinned host memory). At N>1 each rank does the same with its shard.
    e2e = None
    if not args.no_e2e and nlocal <= 31:     # 2 x 16 GiB of pinned host memory per rank at 30 qubits per GPU
        nbytes = 16 << nlocal
        del state
        torch.cuda.empty_cache()
        try:
            host_in = torch.zeros(1 << nlocal, dtype=torch.complex128).pin_memory()
            if rank == 0:
                host_in[0] = 1.0
            host_out = torch.empty(1 << nlocal, dtype=torch.complex128).pin_memory()
            pinned_ok = 1
        except (RuntimeError, MemoryError):
            host_in = host_out = None
            pinned_ok = 0
        if world > 1:       # every rank must take the same branch: the e2e loop contains collectives
            flag = torch.tensor([pinned_ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            pinned_ok = int(flag.item())
    if not args.no_e2e and nlocal <= 31 and not pinned_ok:
        e2e = {'value': None, 'unit': 'gates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
               'skipped': 'pinned host buffers (2 x {} GiB per rank) could not be allocated'.format(nbytes >> 30)}
    elif not args.no_e2e and nlocal <= 31:
        # N=1: Circuit.run_pipelined streams the states through upload / sweeps / download on three streams
        # (every step's 16 GiB input and 16 GiB result cross PCIe inside the timed region; pipeline fill and
        # drain are inside it too). N>1: each rank uploads its shard, runs the sharded circuit, downloads it.
        e2e_steps = max(args.steps, 16) if runner is None else max(1, min(args.steps, 3))

        def e2e_steps_run(count):
            if runner is None:
                circ.run_pipelined([host_in] * count, [host_out] * count, depth=3)
            else:
                for _ in range(count):
                    dstate = runner.execute(host_in.to(dev, non_blocking=False))
                    host_out.copy_(dstate, non_blocking=False)

        e2e_steps_run(1)      # warm-up (the first call also builds and uploads the plan)
        barrier()
        tw0 = time.perf_counter()
        e2e_steps_run(e2e_steps)
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - tw0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {'value': ngates / (e2e_ms * 1e-3) * world, 'unit': 'gates/s', 'h2d_bytes_per_step': nbytes * world,
               'd2h_bytes_per_step': nbytes * world, 'ms_per_step': e2e_ms, 'steps': e2e_steps,
               'api': ('Circuit.run_pipelined(pinned host states -> pinned host results), 3 device buffers, upload / '
                       'sweeps / download overlapped across consecutive steps') if runner is None else
                      'pinned host shard -> ShardedCircuit.execute -> pinned host shard'}
        del host_in, host_out

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_c_port(nq, args.depth, args.seed, args.cpu_seconds)
        except MemoryError:
            cpu = {'value': None, 'unit': 'gates/s', 'cores': 0, 'kind': 'port', 'sample': 'host out of memory'}

    if rank == 0:
        line = {
            'metric': 'gates/s', 'value': value, 'unit': 'gates/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'complex128', 'data': 'synthetic',
            'config': workload_config(nq, args.depth, args.seed, world, ngates),
            'plan': dict(stats, plan_seconds=plan_seconds, tile_bits=args.tile_bits or planner.DEFAULT_TILE_BITS),
            'circuit_gates_per_s': gates_per_s,
            'unfused_equivalent_gbs': ngates * algo_bytes / (ms_step * 1e-3) / 1e9,
            'norm_error_after_run': norm_err,
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches),
            'clocks': clocks,
        }
        if runner is not None:
            line['comm'] = runner.comm_summary()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
