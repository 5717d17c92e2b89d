#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: gates/s (and achieved HBM GB/s) of a depth-20 layered random
1q/2q circuit on a 30-qubit complex128 state per GPU (workload W-B of SURVEY.md 8d, config "Random circuit 30
qubits complex128 on 1xB200 with gate fusion"); with --gpus N the state has 30+log2(N) qubits, sharded on its
top qubits (weak scaling, BASELINE.json configs[4] shape).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--qubits Q] [--depth D] [--impl reference]

One step = one pass of the whole circuit over the state. Prints ONE JSON line (rank 0).
  value        device-timed (CUDA events on the launching stream, max over ranks), state resident in HBM
  e2e          the same circuit with HOST buffers, host<->device copies inside the timed region: one GPU =
               Circuit.run_pipelined (pinned host states -> upload / sweeps / download on three streams -> pinned host
               results); sharded = each rank's shard from / to host memory around ShardedCircuit.execute
  e2e_reference_shaped   the reference's own call sequence, un-pipelined: qf.State(host numpy array) ->
               Circuit.run -> qf.asarray (one GPU)
  e2e_cold     first Circuit.run of a fresh process: planning + kernel generation / compilation + execution
  roofline     dominant kernel = the sweep kernel (one launch = one read + one write of the state = 32 B/amplitude)
  cpu_baseline the C/OpenMP restatement of the reference's tensormul on the host cores, bounded sample
  parity_max_abs (N > 1)  sharded path against the single-GPU engine on a 21-qubit-per-GPU circuit, inside the run
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
    ap.add_argument('--qubits', type=int, default=None,
                    help='qubits per GPU (default 30 on one GPU, 33 when sharded: BASELINE.json configs[3] / [4])')
    ap.add_argument('--config', default='headline', choices=['headline', 'c1', 'c2', 'c3'],
                    help='headline = the W-B circuit of BASELINE.json metric; c1 / c2 / c3 = the other configurations '
                         '(20-qubit circuit, QAOA gradient step, 14-qubit density evolution), one JSON line each')
    ap.add_argument('--no-parity', action='store_true', help='skip the sharded-vs-single-GPU parity run (N > 1)')
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
    """(dram read+write bytes per sweep launch, where the figure comes from): the committed `ncu --set full` capture
    of this round's kernel on the same workload (profiles/sweep_traffic.json names the report), or (None, why)."""
    path = os.path.join(ROOT, 'profiles', 'sweep_traffic.json')
    try:
        with open(path) as f:
            rec = json.load(f)
        if int(rec.get('nbits', -1)) == nbits:
            return float(rec['dram_bytes_per_launch']), rec.get('source', 'profiles/sweep_traffic.json')
    except Exception:
        pass
    return None, 'no ncu capture for {} index bits per GPU'.format(nbits)


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


REFERENCE_PROBE_QUBITS = 26      # the reference arm's bounded sample runs at this size (1 GiB state, ~1 s per gate)


def representative_gates(m: int):
    """1-qubit / 2-qubit gates on low / middle / high axes: np.einsum's cost depends on the
    axis, so the sample cycles through all of them."""
    lo, mid, hi = 1, m // 2, m - 2
    return [('H', (), (hi,)), ('CNOT', (), (lo, mid)), ('RX', (0.7,), (mid,)), ('CZ', (), (mid, hi)),
            ('RY', (0.9,), (lo,)), ('CNOT', (), (hi, lo)), ('T', (), (mid,)), ('CNOT', (), (mid + 1, mid))]


def reference_einsum_step(specs, start: int, ngates: int, state):
    """`ngates` gates with the reference's own np.einsum call (numpybk.py:159-214), cycling through `specs`."""
    from oracle import qf_oracle as O
    t0 = time.perf_counter()
    for i in range(ngates):
        name, params, qubits = specs[(start + i) % len(specs)]
        state = O.tensormul(O.as_tensor(O.gate_matrix(name, params)), state, list(qubits))
    return time.perf_counter() - t0, state


def run_reference_arm(args):
    """bench.py --impl reference: the reference's algorithm on the host (single-threaded np.einsum, the reference's
    own subscripts). A full 30-qubit gate takes ~25 s, so every step is a bounded sample: a few representative
    gates (1q / 2q x low / mid / high axis) on a 2^26-amplitude state, scaled to the benchmark's size by the
    measured linear cost of einsum in the number of amplitudes (labelled as extrapolated in `sample`). Sized so
    that the whole --steps / --warmup run ends within a few minutes whatever the number of GPUs of the main arm."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import numpy as np
    from quantumflow_b200 import workloads
    world = max(1, args.gpus)
    p = world.bit_length() - 1
    nq = default_local_qubits(args, world) + p
    ngates = workloads.wb_gate_count(nq, args.depth)
    m = min(nq, REFERENCE_PROBE_QUBITS)
    specs = representative_gates(m)
    state = np.zeros([2] * m, dtype=np.complex128)
    state[(0,) * m] = 1
    dt_probe, state = reference_einsum_step(specs, 0, 2, state)
    budget = 120.0 / max(1, args.steps + args.warmup)              # seconds per step
    per_step = max(1, min(len(specs), int(budget / max(dt_probe / 2, 1e-6))))
    pos = 2
    for _ in range(args.warmup):
        _, state = reference_einsum_step(specs, pos, per_step, state)
        pos += per_step
    total = 0.0
    for _ in range(args.steps):
        dt, state = reference_einsum_step(specs, pos, per_step, state)
        pos += per_step
        total += dt
    rate_m = per_step * args.steps / total                          # gates/s on 2^m amplitudes
    gates_per_s = rate_m / 2.0 ** (nq - m)                          # extrapolated to the benchmark's 2^nq amplitudes
    value = gates_per_s * 2.0 ** (nq - 30)                          # the main arm's unit (30-qubit equivalents)
    sample = ('{} representative gates per step (H / RX / RY / T / CNOT / CZ on low, middle and high axes) on a '
              '{}-qubit state with the reference subscripts of np.einsum, single thread, {} steps: {:.3f} gates/s '
              'measured at {} qubits, EXTRAPOLATED x 2^-{} to {} qubits (einsum cost is linear in the number of '
              'amplitudes; a full {}-qubit run of the {} gates would take {:.1f} h)'.format(
                  per_step, m, args.steps, rate_m, m, nq - m, nq, nq, ngates, ngates / gates_per_s / 3600.0))
    line = {
        'impl': 'reference', 'metric': 'gates/s', 'value': value, 'unit': value_unit(world), 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex128',
        'data': 'synthetic',
        'config': workload_config(nq, args.depth, args.seed, world, ngates),
        'circuit_gates_per_s': gates_per_s,
        'cpu_baseline': {'value': value, 'unit': value_unit(world), 'cores': 1, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': value_unit(world), 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def default_local_qubits(args, world: int) -> int:
    """Qubits per GPU: BASELINE.json configs[3] (30 qubits) on one GPU, configs[4] (33 per GPU: 34 / 35 / 36 qubits on
    2 / 4 / 8 GPUs) when sharded."""
    if args.qubits:
        return args.qubits
    return 30 if world == 1 else 33


def value_unit(world: int) -> str:
    return 'gates/s' if world == 1 else 'gates/s in 30-qubit equivalents (circuit gates/s x 2^(qubits-30))'


def workload_config(nq, depth, seed, world, ngates):
    return {'workload': 'W-B depth-{} layered random 1q/2q circuit ({{H,X,T,RX,RY,RZ}} per qubit + CNOT/CZ on a '
                        'random perfect matching per layer), {} qubits, complex128, seed {}, {} gates'.format(
                            depth, nq, seed, ngates),
            'qubits': nq, 'depth': depth, 'gates': ngates, 'seed': seed,
            'sharding': 'none' if world == 1 else 'top {} qubits over {} GPUs'.format(world.bit_length() - 1, world),
            'unit_of_work': 'one gate of the circuit applied to 2^30 amplitudes: value = circuit gates/s x 2^(qubits-30), '
                            'plain gates/s for the 30-qubit circuit on one GPU; a sharded run of q qubits does '
                            '2^(q-30) such units per gate, so value / (n_gpus x one-GPU value) is the weak-scaling '
                            'efficiency; the raw rate is circuit_gates_per_s',
            'l2': 'state is {} MiB per GPU, far larger than the 126 MB L2 (no flush needed)'.format(
                (16 << (nq - (world.bit_length() - 1))) >> 20)}


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------

def sharded_parity(qf, world: int, rank: int, dev) -> float:
    """Multi-GPU parity inside the bench (no test can skip it): a 21-qubit-per-GPU W-B circuit through the sharded
    path (sweep-specialised kernels forced on, 1 MiB staging chunks on the NCCL path) against the single-GPU engine
    running the whole circuit on this rank's device with the interpreter. max-abs amplitude difference over all
    ranks (tolerance 1e-10, BASELINE.json north_star)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from quantumflow_b200 import sharded, workloads
    p = world.bit_length() - 1
    n = 21 + p
    saved = {k: os.environ.get(k) for k in ('QFB_JIT', 'QFB_REG_BITS')}
    try:
        os.environ['QFB_JIT'] = '1'
        circ = workloads.wb_circuit(qf, n, 8, 11)
        runner = sharded.ShardedCircuit(circ, n, world, rank, staging_bytes=1 << 20)
        shard = torch.zeros(1 << runner.nl, dtype=torch.complex128, device=dev)
        if rank == 0:
            shard[0] = 1.0
        shard = runner.execute(shard)
        os.environ['QFB_JIT'] = '0'
        full = qf.asarray(workloads.wb_circuit(qf, n, 8, 11).run().tensor).reshape(-1)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    local = np.arange(1 << runner.nl, dtype=np.int64) | (np.int64(rank) << runner.nl)
    logical = np.zeros_like(local)
    for b, pos in enumerate(runner.final_phys_of):
        logical |= ((local >> pos) & 1) << b
    err = float(np.abs(shard.cpu().numpy() - full[logical]).max())
    t = torch.tensor([err], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def spot_amplitudes(state, nlocal: int, nq: int, rank: int, world: int, phys_of, dev):
    """16 amplitudes at fixed logical indices (a 1-GPU run of the same circuit and seed reproduces them where the
    state fits one GPU): [[logical index, re, im], ...]."""
    import torch
    import torch.distributed as dist
    out = torch.zeros(16, 2, dtype=torch.float64, device=dev)
    indices = [(i * 0x9E3779B97F4A7C15 + 12345) % (1 << nq) for i in range(16)]
    for k, logical in enumerate(indices):
        physical = 0
        for b in range(nq):
            physical |= ((logical >> b) & 1) << (phys_of[b] if phys_of is not None else b)
        if physical >> nlocal == rank:
            amp = state[physical & ((1 << nlocal) - 1)]
            out[k, 0], out[k, 1] = amp.real, amp.imag
    if world > 1:
        dist.all_reduce(out)
    vals = out.cpu().tolist()
    return [[int(i), v[0], v[1]] for i, v in zip(indices, vals)]


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
        return
    if args.config != 'headline':
        run_config(args)
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
    dev = torch.device('cuda', local_rank)
    torch.zeros(1, device=dev)                      # CUDA context (not part of the cold-start figure)
    torch.cuda.synchronize()

    import quantumflow_b200 as qf
    from quantumflow_b200 import engine, planner, workloads

    p = world.bit_length() - 1
    assert (1 << p) == world, 'number of GPUs must be a power of two'
    nlocal = default_local_qubits(args, world)
    nq = nlocal + p
    specs = workloads.wb_gate_list(nq, args.depth, args.seed)
    ngates = len(specs)
    scale30 = 2.0 ** (nq - 30)                      # 30-qubit equivalents per gate of this circuit

    # ---- cold start: circuit -> plan -> sweep-specialised kernels (PTX generation + compilation) -> first run ----
    t_cold0 = time.perf_counter()
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
    plan_seconds = time.perf_counter() - t_cold0
    t_jit0 = time.perf_counter()
    for seg in segments:
        if seg.kind == 'plan' and seg.uploaded is None:
            seg.uploaded = engine.UploadedPlan(seg.blob)     # validates, compiles (cached per structure), loads
    jit_seconds = time.perf_counter() - t_jit0
    stats = planner.plan_stats(segments)
    import hashlib
    stats['plan_sha256'] = hashlib.sha256(b''.join(s.blob for s in segments if s.blob)).hexdigest()[:16]
    stats['reg_bits'] = planner.default_reg_bits(nlocal)

    state = torch.zeros(1 << nlocal, dtype=torch.complex128, device=dev)
    if rank == 0:
        state[0] = 1.0

    def step():
        nonlocal state
        if runner is None:
            qf.Circuit._execute(segments, state)
        else:
            state = runner.execute(state)     # in place (peer-memory block exchange at the remaps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()
    barrier()
    cold_seconds = time.perf_counter() - t_cold0
    spots = spot_amplitudes(state, nlocal, nq, rank, world, runner.final_phys_of if runner is not None else None, dev)
    cold = {'plan_seconds': plan_seconds, 'kernel_build_seconds': jit_seconds, 'total_seconds': cold_seconds,
            'value': ngates / cold_seconds * scale30, 'unit': value_unit(world),
            'what': 'first Circuit.run of a fresh process: build the circuit, plan it, generate + compile + load the '
                    'sweep-specialised kernels, allocate |0..0>, run once (CUDA context creation excluded)'}

    # ---- parity of the sharded path (N > 1), inside the run the driver records ----
    parity = None
    if world > 1 and not args.no_parity:
        parity = sharded_parity(qf, world, rank, dev)

    for _ in range(max(0, args.warmup - 1)):
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
    value = gates_per_s * scale30

    # unitarity check after all steps: the state must still be normalised (work was really done, correctly)
    n2 = engine.norm2(state)
    if world > 1:
        dist.all_reduce(n2, op=dist.ReduceOp.SUM)
    norm_err = abs(float(n2) - 1.0)

    # roofline of the dominant kernel: sweep launches only (remap / fallback kernels are excluded from the count)
    nsweeps = stats['sweeps']
    comm_ms = runner.comm_ms_per_step() if runner is not None else 0.0
    pipelined = runner is not None and runner.comm_summary().get('pipelined_remaps_per_step', 0) > 0
    # exchange kernels of a pipelined remap run BESIDE sweeps, so their time cannot be subtracted from the step: the
    # sharded figure is then the conservative one (algorithmic sweep bytes / whole step, exchange included)
    sweep_ms = (ms_step - (0.0 if pipelined else comm_ms)) / max(1, nsweeps)
    algo_bytes = 32.0 * (1 << nlocal)
    peak, peak_src = measured_peak_gbs()
    achieved = algo_bytes / (sweep_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic_bytes(nlocal)
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': traffic_src,
                'kernel': 'qfb_sweep (sweep-specialised, {} register bits, tile 2^{})'.format(
                    stats['reg_bits'], args.tile_bits or planner.default_tile_bits(stats['reg_bits']))
                if os.environ.get('QFB_JIT', '1') != '0' else 'sweep_kernel<{}> (interpreter)'.format(
                    args.tile_bits or planner.default_tile_bits(stats['reg_bits'])),
                'algorithmic_bytes_per_launch': algo_bytes, 'launches_per_step': nsweeps,
                'avg_launch_ms': sweep_ms, 'peak_source': peak_src}
    if pipelined:
        roofline['note'] = ('sharded run with pipelined remaps: avg_launch_ms = whole step / sweep launches (the exchange '
                            'kernels overlap the sweeps and are not subtracted), so frac is a lower bound of the sweep '
                            "kernel's own fraction; the one-GPU line carries the kernel's roofline")

    # ---- end to end with HOST buffers ----
    e2e = None
    e2e_ref_shaped = None
    nbytes = 16 << nlocal
    if not args.no_e2e:
        if nlocal <= 31:
            del state
            torch.cuda.empty_cache()
            e2e, e2e_ref_shaped = e2e_resident_host(args, qf, circ, runner, nlocal, nq, ngates, scale30, world, rank,
                                                    dev, barrier)
        else:
            # the shard buffer is reused (it is mapped into the peers' address spaces for the remaps)
            e2e = e2e_streamed(args, runner, state, nlocal, ngates, scale30, world, rank, dev, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_c_port(nq, args.depth, args.seed, args.cpu_seconds)
        except MemoryError:
            cpu = {'value': None, 'unit': 'gates/s', 'cores': 0, 'kind': 'port', 'sample': 'host out of memory'}

    if rank == 0:
        line = {
            'metric': 'gates/s', 'value': value, 'unit': value_unit(world), 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'complex128', 'data': 'synthetic',
            'config': workload_config(nq, args.depth, args.seed, world, ngates),
            'plan': dict(stats, plan_seconds=plan_seconds, kernel_build_seconds=jit_seconds,
                         tile_bits=args.tile_bits or planner.default_tile_bits(stats['reg_bits'])),
            'circuit_gates_per_s': gates_per_s,
            'unfused_equivalent_gbs': ngates * algo_bytes / (ms_step * 1e-3) / 1e9,
            'norm_error_after_run': norm_err, 'parity_max_abs': parity, 'spot_amplitudes': spots,
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'e2e_reference_shaped': e2e_ref_shaped,
            'e2e_cold': cold, 'gpu_launches': int(launches), 'clocks': clocks,
        }
        if runner is not None:
            line['comm'] = runner.comm_summary()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_config(args):
    """bench.py --config c1|c2|c3: the BASELINE.json configurations that are parity cases rather than the headline
    (SURVEY 8d W-B 20 qubits, W-Q QAOA gradient step, W-D 14-qubit density evolution), through the public API on
    one GPU, one JSON line. Timed with CUDA events around `steps` calls after `warmup` calls; the state is created
    inside the timed call as the reference's Circuit.run() does."""
    import numpy as np
    import torch
    torch.cuda.set_device(0)
    import quantumflow_b200 as qf
    from quantumflow_b200 import engine, workloads
    peak, peak_src = measured_peak_gbs()
    line = None
    if args.config == 'c1':
        n = args.qubits or 20
        circ = workloads.wb_circuit(qf, n, args.depth, args.seed)
        units, unit, what = len(circ.elements), 'gates/s', 'Circuit.run() of the W-B circuit, {} qubits depth {}, {} gates'.format(n, args.depth, len(circ.elements))
        call = lambda: circ.run()                                                       # noqa: E731
        state_bytes = 16 << n
    elif args.config == 'c3':
        n = args.qubits or 14
        circ = workloads.wd_circuit(qf, n, args.depth, args.seed, kraus=True)
        units, unit, what = len(circ.elements), 'ops/s', ('Circuit.evolve() of the W-D density workload, {} qubits depth {}: '
                                                         '{} operations (RX, CNOT, Depolarizing(0.01) as Kraus), rho = {} MiB'
                                                         .format(n, args.depth, len(circ.elements), (16 << (2 * n)) >> 20))
        call = lambda: circ.evolve()                                                    # noqa: E731
        state_bytes = 16 << (2 * n)
    else:
        import networkx as nx
        n, steps_p = args.qubits or 6, 5
        graph = nx.gnp_random_graph(n, 0.5, seed=args.seed)
        cuts = qf.graph_cuts(graph)
        np.random.seed(args.seed)
        beta = torch.tensor(np.random.normal(0.5, 0.01, size=steps_p), requires_grad=True)
        gamma = torch.tensor(np.random.normal(0.5, 0.01, size=steps_p), requires_grad=True)

        def call():
            circ = qf.qubo_circuit(graph, steps_p, beta, gamma)
            expect = circ.run().expectation(cuts)
            (-expect).backward()
            with torch.no_grad():
                beta.sub_(0.01 * beta.grad)
                gamma.sub_(0.01 * gamma.grad)
            beta.grad = None
            gamma.grad = None
            return expect
        units, unit = 1, 'gradient steps/s'
        what = ('QAOA MaxCut gradient step (examples/qaoa_maxcut.py): gnp_random_graph({}, 0.5, seed {}) with {} edges, '
                '{} QAOA steps, forward + backward through the torch.autograd bridge + plain gradient descent'
                .format(n, args.seed, graph.number_of_edges(), steps_p))
        state_bytes = 16 << n
    for _ in range(max(1, args.warmup)):
        call()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    launches0 = engine.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    tw0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        call()
    ev1.record()
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - tw0) / args.steps
    t1 = time.time()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = (engine.launch_count() - launches0) / args.steps
    clocks = sampler.stop(t0, t1)
    line = {'metric': unit, 'value': units / (ms * 1e-3), 'unit': unit, 'n_gpus': 1, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'wall_ms_per_step': wall_ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex128', 'data': 'synthetic',
            'config': {'workload': what, 'name': args.config},
            'gpu_launches': launches * args.steps, 'launches_per_step': launches, 'clocks': clocks}
    if args.config == 'c3':
        segs = [seg for cache in [circ.__dict__.get('_plan_cache', {})] for hit in cache.values() for seg in hit[0]]
        nsweeps = sum(seg.nsweeps for seg in segs)
        algo = 2.0 * state_bytes
        line['roofline'] = {'bound': 'hbm', 'achieved': nsweeps * algo / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                            'frac': nsweeps * algo / (ms * 1e-3) / 1e9 / peak, 'traffic': None,
                            'algorithmic_bytes_per_launch': algo, 'launches_per_step': nsweeps,
                            'avg_launch_ms': ms / max(1, nsweeps), 'peak_source': peak_src,
                            'note': 'every sweep of rho carries dense 2-bit superoperators (16 or 4 FP64 instructions per '
                                    'amplitude each): the FP64 pipe, not HBM, bounds these sweeps (DESIGN.md section 4)'}
    else:
        line['roofline'] = {'bound': 'launch latency', 'achieved': None, 'peak': None, 'unit': 'GB/s', 'frac': None,
                            'traffic': None,
                            'note': 'state of {} bytes is L2 / cache resident: the step is bound by kernel launches and '
                                    'host-side work, not by HBM (SURVEY 8d)'.format(state_bytes)}
    print(json.dumps(line))


def e2e_resident_host(args, qf, circ, runner, nlocal, nq, ngates, scale30, world, rank, dev, barrier):
    """States that fit pinned host memory (<= 31 qubits per GPU). N=1: Circuit.run_pipelined streams host states
    through upload / sweeps / download on three streams, plus the reference-shaped sequence State(host array) ->
    Circuit.run -> asarray. N>1: each rank uploads its shard, runs the sharded circuit, downloads it."""
    import numpy as np
    import torch
    import torch.distributed as dist
    nbytes = 16 << nlocal
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
    if not pinned_ok:
        return ({'value': None, 'unit': value_unit(world), 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
                 'skipped': 'pinned host buffers (2 x {} GiB per rank) could not be allocated'.format(nbytes >> 30)},
                None)
    e2e_steps = max(args.steps, 16) if runner is None else max(1, min(args.steps, 3))

    def e2e_steps_run(count):
        if runner is None:
            circ.run_pipelined([host_in] * count, [host_out] * count, depth=3)
        else:
            for _ in range(count):
                dstate = runner.execute(host_in.to(dev, non_blocking=False))
                host_out.copy_(dstate, non_blocking=False)

    e2e_steps_run(1)      # warm-up
    barrier()
    tw0 = time.perf_counter()
    e2e_steps_run(e2e_steps)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - tw0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {'value': ngates / (e2e_ms * 1e-3) * scale30, 'unit': value_unit(world),
           'h2d_bytes_per_step': nbytes * world, 'd2h_bytes_per_step': nbytes * world, 'ms_per_step': e2e_ms,
           'steps': e2e_steps,
           'api': ('Circuit.run_pipelined(pinned host states -> pinned host results), 3 device buffers, upload / '
                   'sweeps / download overlapped across consecutive steps') if runner is None else
                  'pinned host shard -> ShardedCircuit.execute -> pinned host shard'}
    ref_shaped = None
    if runner is None:
        # the reference's own call sequence (qf.State(array) -> Circuit.run -> asarray), pageable numpy arrays
        del host_out
        host_np = host_in.numpy().reshape([2] * nq)
        times = []
        for _ in range(2):
            torch.cuda.synchronize()
            tr0 = time.perf_counter()
            ket = qf.State(host_np)
            out = circ.run(ket)
            res = qf.asarray(out.tensor)
            times.append(time.perf_counter() - tr0)
            assert abs(float(np.vdot(res[(0,) * nq], res[(0,) * nq]).real)) >= 0.0
            del ket, out, res
        ref_ms = 1e3 * min(times)
        ref_shaped = {'value': ngates / (ref_ms * 1e-3), 'unit': 'gates/s', 'ms_per_step': ref_ms,
                      'h2d_bytes_per_step': nbytes, 'd2h_bytes_per_step': nbytes,
                      'api': 'qf.State(host numpy array) -> Circuit.run(ket) -> qf.asarray(result): the reference\'s '
                             'call sequence, un-pipelined (upload, clone of the caller\'s state, sweeps, download in a row)'}
    return e2e, ref_shaped


def e2e_streamed(args, runner, shard, nlocal, ngates, scale30, world, rank, dev, barrier):
    """Shards that do not fit pinned host memory (33 qubits per GPU = 128 GiB per rank): the shard crosses PCIe in
    1 GiB pieces through two pinned staging buffers, host -> device before the circuit and device -> host after it,
    inside the timed region. The host never holds the whole shard: the input pieces are produced in the staging
    buffer (|0..0>: zeros, amplitude 1 in rank 0's first piece), the result pieces are overwritten after a checksum."""
    import torch
    import torch.distributed as dist
    nbytes = 16 << nlocal
    piece = 1 << 26                                    # amplitudes per piece (1 GiB)
    npieces = (1 << nlocal) // piece
    stage = [torch.zeros(piece, dtype=torch.complex128).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)

    def one_step():
        nonlocal shard
        with torch.cuda.stream(copy_stream):
            for i in range(npieces):
                buf = stage[i % 2]
                if i < 2:
                    buf.zero_()
                    if rank == 0 and i == 0:
                        buf[0] = 1.0
                elif rank == 0 and i == 2:
                    copy_stream.synchronize()          # piece 0 (the one that holds amplitude 1) has left stage[0]
                    stage[0][0] = 0.0
                shard[i * piece:(i + 1) * piece].copy_(buf, non_blocking=True)
        copy_stream.synchronize()
        shard = runner.execute(shard)
        torch.cuda.synchronize()
        acc = 0.0
        with torch.cuda.stream(copy_stream):
            for i in range(npieces):
                buf = stage[i % 2]
                buf.copy_(shard[i * piece:(i + 1) * piece], non_blocking=True)
                if i % 2 == 1:
                    copy_stream.synchronize()
                    acc += float(stage[0][0].real) + float(stage[1][0].real)
        copy_stream.synchronize()
        return acc

    steps = 1 if args.steps < 4 else 2
    barrier()
    tw0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - tw0) / steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    return {'value': ngates / (e2e_ms * 1e-3) * scale30, 'unit': value_unit(world),
            'h2d_bytes_per_step': nbytes * world, 'd2h_bytes_per_step': nbytes * world, 'ms_per_step': e2e_ms,
            'steps': steps,
            'api': 'each rank: host pieces of 1 GiB (two pinned staging buffers) -> shard -> ShardedCircuit.execute -> '
                   'host pieces; {} GiB up and {} GiB down per rank and step inside the timed region; the host '
                   'never holds a whole shard'.format(nbytes >> 30, nbytes >> 30)}


if __name__ == '__main__':
    main()
