"""State vectors sharded over P = 2^p GPUs of one box (one process per GPU, torch.distributed).

The reference has no distributed path; this is the B200-native scaling axis. Flat index =
(rank bits | local bits): rank r owns the contiguous slice [r 2^nl, (r+1) 2^nl). A logical->physical bit map is
kept on the host, so "swapping back" is never needed.

* diagonal operators and control qubits on rank bits cost no communication: the kernels resolve them from
  `index_hi` (= the rank), see qfb_apply_diag / qfb_run_plan;
* an operator that MIXES a rank bit triggers a remap: k rank bits are exchanged with the k top local bits by a
  pairwise block exchange (for k = p this is the all-to-all of SURVEY 8e). The outgoing logical qubits are first moved
  to the top local positions by an in-place bit permutation that rides on the final store of the stage's last sweep
  (planner.attach_permutation: no extra pass over the shard);
* the exchange is IN PLACE: block j of the shard trades places with block `mine` of the peer whose selected rank bits
  equal j. On GPUs it is ONE kernel per remap and rank over NVLink peer memory (csrc/qfb_remap.cu: the peers' shards
  are mapped through CUDA IPC, of every pair's block the lower rank swaps the first half and the higher rank the
  second half; no staging buffer, no NCCL call on the data path), and it is PIPELINED slice by slice with the sweeps on
  either side of it (`_remap_pipelined`: sweeps of slice s+1 | exchange of slice s | sweeps of slice s-1 on two
  streams, ranks ordered by a barrier kernel through peer memory). Memory = the shard alone, which is what lets a
  2^33-amplitude (128 GiB) shard live on a 180 GB GPU. The round-1 path (`_exchange`: XOR-paired isend / irecv chunks
  through two staging buffers) is what the gloo tests run and what QFB_REMAP=nccl selects;
* which qubits become global is decided Belady-style: the ones whose next mixing use is farthest away.

`ShardedCircuit` only needs three callables for the local work (run segments, permute bits, allocate scratch),
so the scheduling and exchange logic is testable on CPU with gloo and the oracle standing in for the kernels
(tests/test_sharded_cpu.py); on GPU the callables are the libqfb200 paths.
"""
import ctypes
import os
import time
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import classify, planner

BitOp = Tuple[np.ndarray, Sequence[int]]


def _mixing_and_diag_bits(mat: np.ndarray, bits: Sequence[int]) -> Tuple[frozenset, frozenset]:
    """(bits the operator mixes, bits it only reads as control / diagonal)."""
    k = len(bits)
    m = classify.as_matrix(mat, k)
    if classify.is_identity(m):
        return frozenset(), frozenset()
    if classify.is_diagonal(m):
        return frozenset(), frozenset(bits)
    if k == 1:
        return frozenset(bits), frozenset()
    controls, targets, _ = classify.peel_controls(m, k)
    return frozenset(bits[q] for q in targets), frozenset(bits[q] for q in controls)


class Stage:
    """Local work between two remaps: classified operators (planner.POp / planner.Fallback) with PHYSICAL bit
    positions, planned for nl local bits, followed by the in-place local bit permutation `final_perm` (dst local
    bit j <- src local bit final_perm[j]; None = identity) that prepares the next remap."""
    __slots__ = ('items', 'segments', 'final_perm', 'parts')

    def __init__(self, items: List[object], final_perm: Optional[List[int]] = None,
                 parts: Optional[List[Tuple[int, Optional[List[int]]]]] = None):
        self.items = items
        self.segments = None
        self.final_perm = final_perm
        # the sweeps the scheduler formed: (number of items, physical tile bits | None for a Fallback) in order;
        # the stage planner keeps this split (planner.build_segments_from_items, preset=)
        self.parts = parts

    @property
    def bitops(self) -> List[BitOp]:
        """The stage as (matrix, physical bits) operators, for reference executors (tests)."""
        return [planner.item_bitop(it) for it in self.items]


class Remap:
    """Exchange rank bits `rank_positions` (positions inside the rank index) with the top-k local bits."""
    __slots__ = ('rank_positions',)

    def __init__(self, rank_positions: List[int]):
        self.rank_positions = rank_positions


# A sweep with less than this fraction of the work of the plan's average sweep so far is "thin": when operators
# are waiting for a remap, the stage ends instead of spending a whole pass over the shard on it. schedule() tries
# these thresholds and keeps the cheapest schedule under the cost model below.
THIN_CANDIDATES = (0.0, 0.25, 0.4, 0.6)
# cost of a remap that exchanges k rank bits, in units of the part of a sweep that a saved sweep actually saves
# (its memory pass, ~7 ms at 30 qubits per GPU; the operators' own time moves to other sweeps): the remap moves
# the fraction 1 - 2^-k of the shard each way at ~0.5 TB/s = 15 / 24 / 28 ms for k = 1 / 2 / 3
REMAP_COST_PER_FRACTION = 4.5
# a stage's last sweep is re-formed around the bits the remap moves (saving the bare permutation sweep) when the new
# sweep does at least this fraction of the old one's work; what it leaves behind runs in the next stage
REFORM_MIN_WORK = 0.5
# randomised variants of each sweep's greedy start (fixed seeds): both counts are scheduled, the cost model decides
SWEEP_TRIES_CANDIDATES = (3, 8)


def schedule(nbits: int, p: int, bitops: Sequence[BitOp], tile_bits: int = None, low_bits: int = None,
             max_cost: float = None, thin: float = None) -> Tuple[List[object], List[int]]:
    """Split `bitops` into Stage / Remap steps (see _schedule_once). With thin=None the candidates of
    THIN_CANDIDATES are scheduled and the one with the lowest modelled cost (sweeps + remaps in sweep units) is
    returned; the choice is deterministic, so every rank takes the same one."""
    if thin is not None or p == 0:
        steps, phys_of, _ = _schedule_once(nbits, p, bitops, tile_bits, low_bits, max_cost, thin or 0.0,
                                           SWEEP_TRIES_CANDIDATES[0])
        return steps, phys_of
    best = None
    items = planner.classify_all(bitops)       # once for all candidates (the scheduler never changes an operator)
    searches: Dict[tuple, int] = {}            # tile searches shared by the candidates (Planner.search_cache): most repeat
    # (randomised variants per sweep, two-sweep look-ahead of the tile search): the look-ahead runs without
    # variants - picking the variant that does most work NOW is exactly what it is there to avoid
    for tries, lookahead in [(t, False) for t in SWEEP_TRIES_CANDIDATES] + [(0, True)]:
        for cand in (THIN_CANDIDATES[1:3] if lookahead else THIN_CANDIDATES):     # the look-ahead is the slow one
            try:
                steps, phys_of, nsweeps = _schedule_once(nbits, p, bitops, tile_bits, low_bits, max_cost, cand, tries,
                                                         lookahead, items=items, searches=searches)
            except RuntimeError:
                continue          # one candidate that cannot be scheduled must not abort the others
            cost = nsweeps + sum(REMAP_COST_PER_FRACTION * (1.0 - 0.5 ** len(st.rank_positions))
                                 for st in steps if isinstance(st, Remap))
            if best is None or cost < best[0] - 1e-9:
                best = (cost, steps, phys_of)
    if best is None:
        raise RuntimeError('no schedule found')
    return best[1], best[2]


def _schedule_once(nbits: int, p: int, bitops: Sequence[BitOp], tile_bits: int, low_bits: int, max_cost: float,
                   thin_fraction: float, sweep_tries: int = 3, lookahead: bool = False,
                   items: Optional[List[object]] = None, searches: Optional[Dict[tuple, int]] = None
                   ) -> Tuple[List[object], List[int], int]:
    """Split `bitops` (logical bit positions, program order) into Stage / Remap steps.

    Sweeps (passes over the shard) are formed one at a time from the operators that are executable under the
    current logical->physical map - an operator that mixes a global qubit waits, and so does everything that does
    not commute with a waiting operator. A stage ends, and a remap is scheduled, when nothing is executable any
    more OR when the next sweep would be thin while operators are waiting: running every stage to exhaustion ends
    in a tail of a few long dependency chains that costs whole passes for little work (28 passes instead of 20
    on the 33-qubit benchmark). At a remap the p qubits whose next mixing use is farthest away become global
    (Belady). Logical bits 0 .. L-1 never become global: they stay at physical positions 0 .. L-1, the tile
    bits every sweep needs for whole 128-byte lines.

    Returns (steps, final phys_of, number of sweeps formed) where phys_of[logical bit] = physical bit; physical
    bits >= nbits - p are rank bits. Deterministic and identical on every rank."""
    import random
    nl = nbits - p
    phys_of = list(range(nbits))                       # identity: the top p qubits' bits are global
    remaining = planner.classify_all(bitops) if items is None else list(items)
    # tiles are formed here in logical bits and handed to the stage planner (nl local bits): same tile size
    tb = min(planner.default_tile_bits(planner.default_reg_bits(nl)) if tile_bits is None else int(tile_bits), nl)
    pl = planner.Planner(nbits, tb, low_bits, max_cost) if nl >= planner.MIN_TILE_BITS else None
    if pl is not None:
        # without a shared dict: one of its own (the operators of `remaining` stay alive in this call's frame)
        pl.search_cache = searches if searches is not None else {}
    pinned = set(range(pl.L)) if pl is not None else set()

    def mixset(it) -> frozenset:
        return frozenset(it.bits) if isinstance(it, planner.Fallback) else it.mixset

    def anyset(it) -> frozenset:
        return frozenset(it.bits) if isinstance(it, planner.Fallback) else it.anyset

    def diagset(it) -> frozenset:
        return frozenset() if isinstance(it, planner.Fallback) else it.diagset

    def next_sweep(glob: frozenset, remaining: List[object], required: frozenset = frozenset()):
        """(chosen, rest, waiting, tile) - one sweep over the executable operators of `remaining` (tile in logical
        bits, None for a single operator run by the one-gate kernels); waiting = some operator is held back by a
        global qubit. `required`: bits the tile must hold (the positions a remap's local permutation moves)."""
        # a Fallback (dense > 2-bit operator) is a pass of its own: cut the list there
        cut = next((i for i, it in enumerate(remaining) if isinstance(it, planner.Fallback)), len(remaining))
        if cut == 0:
            it = remaining[0]
            if mixset(it) & glob:
                return [], remaining, True, None
            return [it], remaining[1:], False, None
        head, tail = remaining[:cut], remaining[cut:]
        if pl is None:
            it = head[0]
            if mixset(it) & glob:
                return [], remaining, True, None
            return [it], remaining[1:], False, None
        best = None
        # a few randomised variants of the greedy start (fixed seeds; with tile refinement each one is a local
        # search of its own): 8 GPUs 21 -> 19 sweeps, 4 GPUs 20 -> 19, 2 GPUs 18 -> 17 on the benchmark
        for trial in range(1 + (min(pl.tries, sweep_tries if pl.refine else 8) if len(head) >= 64 else 0)):
            rnd = random.Random(trial) if trial else None
            chosen, rest, tile = pl._form_sweep(head, rnd, 1.0 if trial == 0 else 0.9, forbidden=glob,
                                                required=required, lookahead=lookahead and trial == 0)
            if best is None or sum(o.cost for o in chosen) > sum(o.cost for o in best[0]):
                best = (chosen, rest, tile)
        waiting = any(op.kind == 'G' and (op.mixset & glob) for op in best[1]) or \
            (bool(tail) and bool(mixset(tail[0]) & glob))
        return best[0], best[1] + tail, waiting, best[2]

    def take(chosen, tile):
        stage_items.extend(planner.remap_item(op, phys_of) for op in chosen)
        stage_parts.append((len(chosen), None if tile is None else sorted(phys_of[b] for b in tile)))

    steps: List[object] = []
    stage_items: List[object] = []
    stage_parts: List[Tuple[int, Optional[List[int]]]] = []
    costs: List[float] = []              # work of the sweeps scheduled so far (thin-sweep threshold)
    nbare = 0                            # remaps whose local permutation needs a bare sweep of its own
    before_last: Optional[List[object]] = None    # the operator list the stage's last sweep was formed from
    fresh = True                         # no sweep yet since the last remap: take the next one whatever its size
    force = False                        # the remap rule found nothing to exchange: take the thin sweep after all
    while remaining:
        glob = frozenset(b for b in range(nbits) if phys_of[b] >= nl)
        chosen, rest, waiting, tile = next_sweep(glob, remaining)
        work = sum(getattr(o, 'cost', 1.0) for o in chosen)
        thin = bool(costs) and work < thin_fraction * (sum(costs) / len(costs))
        if chosen and not (waiting and thin and not fresh and not force):
            force = False
            before_last = remaining
            take(chosen, tile)
            costs.append(work)
            remaining = rest
            fresh = False
            continue
        if p == 0 or not waiting:
            raise RuntimeError('unschedulable operator')
        # Belady: keep local the bits that are mixed soonest; the p bits used farthest in the future go global
        next_use: Dict[int, int] = {}
        for order, it in enumerate(remaining):
            for b in mixset(it):
                next_use.setdefault(b, order)
        allowed = [b for b in range(nbits) if b not in pinned]
        far = sorted(allowed, key=lambda b: (-next_use.get(b, 1 << 60), -phys_of[b]))
        new_global = set(far[:p])
        cur_global = set(glob)
        outgoing = sorted(new_global - cur_global, key=lambda b: phys_of[b])     # local -> rank
        incoming = sorted(cur_global - new_global, key=lambda b: phys_of[b])     # rank -> local
        k = len(outgoing)
        if k == 0:
            # Belady keeps the current global bits (the waiting operators' bits are used farthest in the future):
            # a remap would change nothing. If the stage was cut short because the next sweep is thin, run that
            # sweep instead (it is what unblocks the order); only a truly stuck schedule is an error.
            if chosen and not force:
                force = True
                continue
            raise RuntimeError('scheduler made no progress')
        # local permutation: outgoing logical bits move to the top-k local positions nl-k .. nl-1
        top = list(range(nl - k, nl))
        logical_at = {phys_of[b]: b for b in range(nbits)}
        src_pos = [phys_of[b] for b in outgoing]
        displaced = [q for q in top if q not in src_pos]             # positions whose content must move down
        holes = [q for q in src_pos if q not in top]                 # positions vacated below the top
        perm = list(range(nl))                                       # dst position j <- src position perm[j]
        stay = [q for q in top if q in src_pos]
        free_top = [q for q in top if q not in stay]
        movers = [q for q in src_pos if q not in top]
        for dst, src in zip(free_top, movers):
            perm[dst] = src
        for dst, src in zip(holes, displaced):
            perm[dst] = src
        local_perm = None if perm == list(range(nl)) else perm
        if local_perm is not None:
            # The permutation is free when the stage's last sweep stores it, i.e. when that sweep's tile holds the
            # moved positions; otherwise it costs a bare pass over the shard (planner.attach_permutation). The
            # sweep was formed before the remap was known, so form it again from the same operator list with the
            # moved bits as required tile bits, and keep the new one if it does at least REFORM_MIN_WORK of the old
            # one's work (the operators it drops stay in the list and run in the next stage).
            moved = frozenset(logical_at[j] for j in range(nl) if perm[j] != j)
            last = stage_parts[-1] if stage_parts else None
            hosted = last is not None and last[1] is not None and all(phys_of[b] in last[1] for b in moved)
            if not hosted and last is not None and last[1] is not None and before_last is not None \
                    and pl is not None and len(moved | pinned) <= pl.M:
                old_items = stage_items[len(stage_items) - last[0]:]
                chosen2, rest2, _waiting2, tile2 = next_sweep(glob, before_last, required=moved)
                old_work = sum(getattr(o, 'cost', 1.0) for o in old_items)
                if tile2 is not None and sum(o.cost for o in chosen2) >= REFORM_MIN_WORK * old_work - 1e-9:
                    del stage_items[len(stage_items) - last[0]:]
                    stage_parts.pop()
                    take(chosen2, tile2)
                    remaining = rest2
                    hosted = True
            if not hosted:
                nbare += 1
            new_logical_at = {j: logical_at[perm[j]] for j in range(nl)}
            for j, b in new_logical_at.items():
                phys_of[b] = j
        # exchange: rank bit position t (physical nl + t) <-> local position top[i], pairing in ascending order
        rank_positions = sorted(phys_of[b] - nl for b in incoming)
        logical_at = {phys_of[b]: b for b in range(nbits)}
        for i, t in enumerate(rank_positions):
            b_local, b_rank = logical_at[top[i]], logical_at[nl + t]
            phys_of[b_local], phys_of[b_rank] = nl + t, top[i]
        if stage_items or local_perm is not None:
            steps.append(Stage(stage_items, local_perm, stage_parts))
            stage_items, stage_parts = [], []
        steps.append(Remap(rank_positions))
        before_last = None
        fresh = True
    if stage_items:
        steps.append(Stage(stage_items, None, stage_parts))
    return steps, phys_of, len(costs) + nbare


class ShardedCircuit:
    """A circuit scheduled for a state sharded over `world` ranks. `execute(shard)` runs it in place on this
    rank's shard (the tensor object may be swapped with the internal scratch buffer: use the returned tensor)."""

    def __init__(self, circuit, nqubits: int, world: int, rank: int, tile_bits: int = None, low_bits: int = None,
                 max_cost: float = None, bitops: Sequence[BitOp] = None,
                 run_stage: Callable = None, permute: Callable = None, group=None,
                 staging_bytes: int = 1 << 30):
        p = world.bit_length() - 1
        assert (1 << p) == world, 'world size must be a power of two'
        self.n, self.p, self.nl = nqubits, p, nqubits - p
        self.world, self.rank, self.group = world, rank, group
        if bitops is None:
            qubits = tuple(range(nqubits)) if circuit is None else tuple(sorted(circuit.qubits))
            assert len(qubits) == nqubits
            bitops = [(g.matrix(), [nqubits - 1 - qubits.index(q) for q in g.qubits]) for g in circuit.elements]
        self._plan_args = dict(tile_bits=tile_bits, low_bits=low_bits, max_cost=max_cost)
        self.steps, self.final_phys_of = schedule(nqubits, p, bitops, **self._plan_args)
        self._run_stage = run_stage or self._run_stage_gpu
        self._permute = permute            # test double only (out of place); the GPU path fuses it into the plan
        self._staging_bytes = int(staging_bytes)
        self._staging = None
        self._peers = None             # peers' shards mapped into this process (CUDA IPC), see _map_peers
        self._peer_key = None
        self._token = None
        self._comm_seconds = 0.0
        self._comm_bytes = 0
        self._remaps = 0
        self._executions = 0
        self._pipelined = 0
        self._pipelined_sweeps = 0
        self._comm_events = []          # (start, end) CUDA events of exchange kernels, resolved lazily (no host sync)
        self._comm_stream = None
        for st in self.steps:
            if isinstance(st, Stage):
                st.segments = planner.build_segments_from_items(self.nl, st.items, final_perm=st.final_perm,
                                                                preset=st.parts,
                                                                **self._plan_args) if run_stage is None else None

    # ---- default (GPU) local work ------------------------------------------------------------------
    def _run_stage_gpu(self, stage: Stage, shard: torch.Tensor) -> None:
        from . import engine
        for seg in stage.segments:
            if seg.kind == 'plan':
                if seg.uploaded is None:
                    seg.uploaded = engine.UploadedPlan(seg.blob)
                seg.uploaded.launch(shard, index_hi=self.rank)
            else:
                engine.apply_operator(shard, seg.mat, seg.bits, inplace=True, index_hi=self.rank)

    # ---- bookkeeping ---------------------------------------------------------------------------------
    def local_segments(self) -> List[planner.Segment]:
        out: List[planner.Segment] = []
        for st in self.steps:
            if isinstance(st, Stage) and st.segments:
                out.extend(st.segments)
        return out

    def _resolve_comm_events(self) -> None:
        if self._comm_events:
            torch.cuda.synchronize()
            for t0, t1 in self._comm_events:
                self._comm_seconds += t0.elapsed_time(t1) * 1e-3
            self._comm_events = []
            if getattr(self, '_barrier_error', None) is not None and int(self._barrier_error.item()) != 0:
                raise RuntimeError('peer barrier {} timed out: a rank did not arrive'.format(
                    int(self._barrier_error.item())))

    def comm_ms_per_step(self) -> float:
        self._resolve_comm_events()
        return 1e3 * self._comm_seconds / max(1, self._executions)

    def comm_summary(self) -> dict:
        ex = max(1, self._executions)
        return {'remaps_per_step': self._remaps / ex, 'bytes_sent_per_rank_per_step': self._comm_bytes / ex,
                'ms_per_step': self.comm_ms_per_step(),
                'path': self._exchange_path,
                'pipelined_remaps_per_step': self._pipelined / ex,
                'sweeps_inside_pipelines_per_step': self._pipelined_sweeps / ex,
                'note': 'in-place pairwise block exchange of k rank bits with the top-k local bits; path "peer": one '
                        'kernel per remap and rank swaps its half of every pair over peer memory (qfb_remap_swap, '
                        'NVLink loads / stores on the IPC-mapped shards, no staging, no NCCL call); path "nccl": '
                        'isend / irecv chunks through two staging buffers. The local bit permutation is fused into '
                        'the last sweep of the preceding stage. Pipelined remaps: the last sweep of the stage, the exchange and the '
                        'first sweep of the next stage run slice by slice (slices = index bits outside both tiles), the exchange '
                        'of slice s beside the sweeps of its neighbours, ranks ordered by a peer-memory barrier; '
                        'ms_per_step is the device time of the exchange kernels (overlapped with sweeps when pipelined)'}

    def reset_comm_counters(self) -> None:
        self._resolve_comm_events()
        self._comm_seconds, self._comm_bytes, self._remaps, self._executions, self._pipelined = 0.0, 0, 0, 0, 0
        self._pipelined_sweeps = 0

    # ---- execution ---------------------------------------------------------------------------------
    _exchange_path = 'nccl'

    def _map_peers(self, shard: torch.Tensor) -> None:
        """Map every peer's shard into this process through CUDA IPC (one process per GPU on one box): afterwards
        self._peers[r] is a device pointer to rank r's shard that this GPU's kernels can load from and store to
        over NVLink. Collective; done once per shard buffer."""
        key = (shard.data_ptr(), shard.numel())
        if self._peer_key == key:
            return
        storage = shard.untyped_storage()
        mine = (storage._share_cuda_(), shard.storage_offset() * shard.element_size())
        gathered = [None] * self.world
        dist.all_gather_object(gathered, mine, group=self.group)
        self._peer_storages, self._peers = [], [0] * self.world
        for r, (handle, offset) in enumerate(gathered):
            if r == self.rank:
                self._peers[r] = shard.data_ptr()
                continue
            # open the peer's allocation with THIS rank's device current: cudaIpcOpenMemHandle then enables peer
            # access from this GPU to the GPU that owns the memory (the tuple's first entry is the owner's device)
            st = torch.UntypedStorage._new_shared_cuda(shard.device.index, *handle[1:])
            self._peer_storages.append(st)                 # keeps the mapping alive
            self._peers[r] = st.data_ptr() + offset
        self._peer_key = key
        self._token = torch.zeros(1, dtype=torch.int32, device=shard.device)
        # flag words of the peer-memory barrier (qfb_peer_barrier), mapped the same way
        self._flags = torch.zeros(64, dtype=torch.int32, device=shard.device)
        self._barrier_error = torch.zeros(1, dtype=torch.int32, device=shard.device)
        torch.cuda.synchronize(shard.device)
        fl = (self._flags.untyped_storage()._share_cuda_(), self._flags.storage_offset() * 4)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, fl, group=self.group)
        self._flag_ptrs = [0] * self.world
        for r, (handle, offset) in enumerate(gathered):
            if r == self.rank:
                self._flag_ptrs[r] = self._flags.data_ptr()
                continue
            st = torch.UntypedStorage._new_shared_cuda(shard.device.index, *handle[1:])
            self._peer_storages.append(st)
            self._flag_ptrs[r] = st.data_ptr() + offset
        self._epoch = 0
        dist.barrier(group=self.group)         # nobody signals before everybody's flags are zero and mapped

    def _rank_barrier(self) -> None:
        """Stream-ordered barrier across the ranks: nobody's later work starts before everybody's earlier work
        (on the current streams) is complete. One 4-byte all-reduce."""
        dist.all_reduce(self._token, group=self.group)

    def _exchange_peer(self, shard: torch.Tensor, rank_positions: List[int]) -> None:
        """The exchange as ONE kernel over peer memory (csrc/qfb_remap.cu): block j of this rank trades places
        with block `mine` of the peer whose selected rank bits equal j; of every pair's block the lower rank swaps
        the first half and the higher rank the second half."""
        from . import _lib, engine
        self._map_peers(shard)
        local, remote, counts = self._swap_runs(shard, rank_positions)
        self._rank_barrier()
        npairs = len(local)
        lib = _lib.load()
        _lib.check(lib.qfb_remap_swap(npairs, (ctypes.c_void_p * npairs)(*local), (ctypes.c_void_p * npairs)(*remote),
                                      (ctypes.c_uint64 * npairs)(*counts), engine._stream()))
        self._rank_barrier()

    def _swap_runs(self, shard: torch.Tensor, rank_positions: List[int]):
        """(local addresses, peer addresses, amplitude counts) of the runs this rank swaps in a remap."""
        k = len(rank_positions)
        blk = shard.numel() >> k
        mine = 0
        for i, t in enumerate(rank_positions):
            mine |= ((self.rank >> t) & 1) << i
        es = shard.element_size()
        local, remote, counts = [], [], []
        for s in range(1, 1 << k):
            j = mine ^ s
            peer = self.rank
            for i, t in enumerate(rank_positions):
                peer = (peer & ~(1 << t)) | (((j >> i) & 1) << t)
            half = blk // 2
            off, n = (0, half) if self.rank < peer else (half, blk - half)
            if n == 0:
                continue
            local.append(shard.data_ptr() + (j * blk + off) * es)
            remote.append(self._peers[peer] + (mine * blk + off) * es)
            counts.append(n)
            self._comm_bytes += n * es * 2        # leaves this GPU: n by its own remote stores, n pulled by the peer
        return local, remote, counts

    def _exchange(self, shard: torch.Tensor, rank_positions: List[int]) -> None:
        """In place: block j of this rank (top-k local bits = j) is swapped with block `mine` of the peer whose
        selected rank bits equal j. Step s pairs mine with mine ^ s on every rank, so both sides of a pair issue
        their transfers in the same order; chunks go through two staging buffers (receive chunk i+1 while chunk
        i is copied into place)."""
        k = len(rank_positions)
        nblocks = 1 << k
        blk = shard.numel() >> k
        mine = 0
        for i, t in enumerate(rank_positions):
            mine |= ((self.rank >> t) & 1) << i
        chunk = max(1, min(blk, self._staging_bytes // shard.element_size()))
        if self._staging is None or self._staging.numel() < 2 * chunk or self._staging.device != shard.device \
                or self._staging.dtype != shard.dtype:
            self._staging = torch.empty(2 * chunk, dtype=shard.dtype, device=shard.device)
        pending = None
        nsent = 0
        for s in range(1, nblocks):
            j = mine ^ s
            peer = self.rank
            for i, t in enumerate(rank_positions):
                peer = (peer & ~(1 << t)) | (((j >> i) & 1) << t)
            for off in range(0, blk, chunk):
                n = min(chunk, blk - off)
                mine_chunk = shard[j * blk + off: j * blk + off + n]
                buf = self._staging[(nsent % 2) * chunk: (nsent % 2) * chunk + n]
                reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, mine_chunk, peer, group=self.group),
                                               dist.P2POp(dist.irecv, buf, peer, group=self.group)])
                nsent += 1
                self._comm_bytes += n * shard.element_size()
                if pending is not None:
                    for req in pending[0]:
                        req.wait()
                    pending[1].copy_(pending[2])
                pending = (reqs, mine_chunk, buf)
        if pending is not None:
            for req in pending[0]:
                req.wait()
            pending[1].copy_(pending[2])

    # ---- pipelined remaps (GPU, peer-memory path) ---------------------------------------------------------
    # A remap sits between the last sweep A of a stage and the first sweep B of the next. Index bits that are outside
    # the tiles of A and of B and below the exchanged block bits cut the shard into 2^v SLICES that never meet in A,
    # in the exchange or in B, so the three are pipelined slice by slice: A runs on slice s+1 while the exchange
    # kernel (another stream, a bounded share of every SM) moves slice s over NVLink and B follows one barrier
    # behind. The ranks are ordered by a barrier through peer memory (qfb_peer_barrier) between the steps of the
    # exchange stream: no host synchronisation, no NCCL kernel that would have to wait for a free SM.
    def _uploaded(self, seg):
        from . import engine
        if seg.uploaded is None:
            seg.uploaded = engine.UploadedPlan(seg.blob)
        return seg.uploaded

    def _pipeline_shape(self, prev: 'Stage', remap: 'Remap', nxt: 'Stage', used_first: int, nxt_has_remap: bool):
        """(selector bits ascending, da, db) of this remap's pipeline, or None when it cannot be pipelined: the last da
        sweeps of `prev` and the first db sweeps of `nxt` run slice by slice around the exchange. Deeper chains hide
        more of the exchange, but every sweep of a chain must keep the selector bits outside its tile. `used_first`: leading sweeps of prev's first plan that the
        previous remap's pipeline has already run."""
        want = int(os.environ.get('QFB_REMAP_SLICE_BITS', '3'))
        depth = int(os.environ.get('QFB_REMAP_CHAIN', '3'))
        if want <= 0 or depth <= 0 or not prev.segments or not nxt.segments:
            return None
        a_seg, b_seg = prev.segments[-1], nxt.segments[0]
        if a_seg.kind != 'plan' or b_seg.kind != 'plan':
            return None
        a, b = self._uploaded(a_seg), self._uploaded(b_seg)
        if not (a.specialised and b.specialised):
            return None
        k = len(remap.rank_positions)
        na, nb = a.nsweeps, b.nsweeps
        da_max = min(depth, na - (used_first if len(prev.segments) == 1 else 0))
        # a one-plan stage that feeds the next remap's pipeline keeps at least half of its sweeps for that one
        db_max = min(depth, nb - (nb + 1) // 2 if (nxt_has_remap and len(nxt.segments) == 1 and nb > 1) else nb)
        if nxt_has_remap and len(nxt.segments) == 1 and nb == 1:
            db_max = 0
        if da_max < 1 or db_max < 1:
            return None
        best = None
        for da in range(1, da_max + 1):
            for db in range(1, db_max + 1):
                common = ~0
                for i in range(na - da, na):
                    common &= a.nontile_mask(i)
                for i in range(db):
                    common &= b.nontile_mask(i)
                # below the half-block split of the exchange; runs of at least 2^12 amplitudes (64 KiB) stay contiguous
                cand = [pos for pos in range(12, self.nl - k - 1) if (common >> pos) & 1]
                nb_bits = min(want, len(cand))
                if nb_bits < 1:
                    continue
                # at least 4 slices first, then the deepest chains (measured at 34 qubits on 2 GPUs: 1 520 / 1 488 /
                # 1 478 ms per step with 1 / 2 / 3 sweeps per side, although the exchange alone is only 1.7 sweeps long:
                # sweeps that run beside the exchange lose HBM bandwidth to it, so more of them fit), then more slices
                key = (min(nb_bits, 2), da + db, nb_bits, -abs(da - db))
                if best is None or key > best[0]:
                    best = (key, cand[-nb_bits:], da, db)
        if best is None:
            return None
        return best[1], best[2], best[3]

    def _peer_barrier(self) -> None:
        from . import _lib, engine
        self._epoch += 1
        ptrs = (ctypes.c_void_p * self.world)(*self._flag_ptrs)
        _lib.check(_lib.load().qfb_peer_barrier(self._flags.data_ptr(), ptrs, self.world, self.rank, self._epoch,
                                                self._barrier_error.data_ptr(), engine._stream()))

    def _run_stage_part(self, stage: 'Stage', shard: torch.Tensor, skip_first: int, skip_last: int) -> None:
        """The stage without the first `skip_first` sweeps of its first plan and the last `skip_last` of its last plan
        (those run inside the pipelines of the neighbouring remaps)."""
        from . import engine
        segs = stage.segments
        for idx, seg in enumerate(segs):
            if seg.kind != 'plan':
                engine.apply_operator(shard, seg.mat, seg.bits, inplace=True, index_hi=self.rank)
                continue
            up = self._uploaded(seg)
            first = skip_first if idx == 0 else 0
            last = up.nsweeps - (skip_last if idx == len(segs) - 1 else 0)
            if last > first:
                up.launch_part(shard, first, last - first, index_hi=self.rank)

    def _remap_pipelined(self, shard: torch.Tensor, prev: 'Stage', remap: 'Remap', nxt: 'Stage',
                         bits: List[int], da: int, db: int) -> None:
        from . import _lib
        lib = _lib.load()
        main = torch.cuda.current_stream(shard.device)
        if getattr(self, '_comm_stream', None) is None:
            self._comm_stream = torch.cuda.Stream(device=shard.device, priority=-1)
        comm = self._comm_stream
        a, b = self._uploaded(prev.segments[-1]), self._uploaded(nxt.segments[0])
        local, remote, counts = self._swap_runs(shard, remap.rank_positions)
        npairs = len(local)
        c_local, c_remote = (ctypes.c_void_p * npairs)(*local), (ctypes.c_void_p * npairs)(*remote)
        c_counts, c_pos = (ctypes.c_uint64 * npairs)(*counts), (ctypes.c_int * len(bits))(*bits)
        mask = sum(1 << pos for pos in bits)
        # the sweeps' CTAs are persistent: a slice launch leaves room for one exchange CTA per SM (QFB_SLICE_ROOM=0: none)
        ctas = int(os.environ.get('QFB_REMAP_CTAS', '1'))
        room = -int(os.environ.get('QFB_SLICE_ROOM', '1'))
        nslices = 1 << len(bits)
        values = [sum(((sl >> t) & 1) << pos for t, pos in enumerate(bits)) for sl in range(nslices)]
        arrived = []
        for sl in range(nslices):
            a.launch_part(shard, a.nsweeps - da, da, index_hi=self.rank, fix_mask=mask, fix_value=values[sl],
                          ctas_per_sm=room)
            ev_a = torch.cuda.Event()
            ev_a.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ev_a)
                # everybody's slice is final (and, from the second slice on, everybody's previous exchange complete)
                self._peer_barrier()
                if sl > 0:
                    ev_x = torch.cuda.Event()
                    ev_x.record(comm)
                    arrived.append(ev_x)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(comm)
                _lib.check(lib.qfb_remap_swap_slice(npairs, c_local, c_remote, c_counts, len(bits), c_pos, values[sl],
                                                    ctas, comm.cuda_stream))
                t1.record(comm)
                self._comm_events.append((t0, t1))
        with torch.cuda.stream(comm):
            self._peer_barrier()
            ev_x = torch.cuda.Event()
            ev_x.record(comm)
            arrived.append(ev_x)
        for sl in range(nslices):
            main.wait_event(arrived[sl])
            b.launch_part(shard, 0, db, index_hi=self.rank, fix_mask=mask, fix_value=values[sl], ctas_per_sm=room)

    def _execute_overlapped(self, shard: torch.Tensor) -> torch.Tensor:
        self._exchange_path = 'peer'
        self._map_peers(shard)
        steps = self.steps
        skip_first = 0
        i = 0
        while i < len(steps):
            st = steps[i]
            if isinstance(st, Remap):           # a remap that no stage's pipeline has taken
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                self._exchange_peer(shard, st.rank_positions)
                t1.record()
                self._comm_events.append((t0, t1))
                self._remaps += 1
                i += 1
                continue
            remap = steps[i + 1] if i + 1 < len(steps) and isinstance(steps[i + 1], Remap) else None
            nxt = steps[i + 2] if remap is not None and i + 2 < len(steps) and isinstance(steps[i + 2], Stage) else None
            shape = None
            if nxt is not None:
                nxt_has_remap = i + 3 < len(steps) and isinstance(steps[i + 3], Remap)
                shape = self._pipeline_shape(st, remap, nxt, skip_first, nxt_has_remap)
            self._run_stage_part(st, shard, skip_first, shape[1] if shape else 0)
            skip_first = 0
            if shape is not None:
                bits, da, db = shape
                self._remap_pipelined(shard, st, remap, nxt, bits, da, db)
                self._pipelined += 1
                self._pipelined_sweeps += da + db
                self._remaps += 1
                skip_first = db
                i += 2
            else:
                i += 1
        self._executions += 1
        return shard

    def execute(self, shard: torch.Tensor) -> torch.Tensor:
        """Runs the circuit in place on this rank's shard and returns it (the same tensor)."""
        assert shard.numel() == 1 << self.nl and shard.is_contiguous()
        timed = shard.is_cuda
        if (timed and self.world > 1 and self._run_stage == self._run_stage_gpu
                and os.environ.get('QFB_REMAP', 'peer') != 'nccl' and os.environ.get('QFB_REMAP_OVERLAP', '1') != '0'):
            return self._execute_overlapped(shard)
        for st in self.steps:
            if isinstance(st, Stage):
                self._run_stage(st, shard)
                if st.final_perm is not None and self._permute is not None:
                    out = torch.empty_like(shard)          # test double: small shards, out of place
                    self._permute(shard, st.final_perm, out)
                    shard.copy_(out)
                continue
            if timed:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            else:
                t0 = time.perf_counter()
            if timed and os.environ.get('QFB_REMAP', 'peer') != 'nccl':
                self._exchange_path = 'peer'
                self._exchange_peer(shard, st.rank_positions)
            else:
                self._exchange(shard, st.rank_positions)
            self._remaps += 1
            if timed:
                ev1.record()
                self._comm_events.append((ev0, ev1))       # resolved when the figures are read: no host sync here
            else:
                self._comm_seconds += time.perf_counter() - t0
        self._executions += 1
        return shard


def gather_logical(shards: Sequence[np.ndarray], nbits: int, p: int, phys_of: Sequence[int]) -> np.ndarray:
    """Test helper: assemble rank shards (physical layout) into the logical flat vector."""
    phys = np.concatenate([np.asarray(s).reshape(-1) for s in shards])
    idx = np.arange(1 << nbits, dtype=np.int64)
    src = np.zeros_like(idx)
    for logical, physical in enumerate(phys_of):
        src |= ((idx >> logical) & 1) << physical
    return phys[src]


# ---------------------------------------------------------------------------------------------------------
# read-out of a sharded state (SURVEY 8e "Reductions / readout"): every rank reduces its own shard with the
# single-GPU kernels (engine.norm2 / marginal / expectation_diag / sample_search) and the partial results -- a
# few doubles -- are combined with one small collective. All functions return the same value on every rank.
# `local=` replaces the device reduction in the CPU tests (gloo, numpy shards).
# ---------------------------------------------------------------------------------------------------------

def _all_sum(values: Sequence[float], like: torch.Tensor, group=None) -> List[float]:
    buf = torch.tensor([float(v) for v in values], dtype=torch.float64, device=like.device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in buf.cpu()]


def scatter_physical(full_logical: np.ndarray, nbits: int, phys_of: Sequence[int]) -> np.ndarray:
    """Test helper, inverse of gather_logical: the physical (rank-major) vector of a logical flat vector."""
    idx = np.arange(1 << nbits, dtype=np.int64)
    dst = np.zeros_like(idx)
    for logical, physical in enumerate(phys_of):
        dst |= ((idx >> logical) & 1) << physical
    phys = np.zeros_like(np.asarray(full_logical).reshape(-1))
    phys[dst] = np.asarray(full_logical).reshape(-1)
    return phys


def norm2(shard: torch.Tensor, group=None, local: Callable = None) -> float:
    """<psi|psi> of the sharded state: local squared norms, one all-reduce of a double."""
    if local is None:
        from . import engine
        mine = float(engine.norm2(shard))
    else:
        mine = float(local(shard))
    return _all_sum([mine], shard, group)[0]


def marginal(shard: torch.Tensor, logical_bit: int, phys_of: Sequence[int], nl: int, rank: int, group=None,
             local: Callable = None, local_norm2: Callable = None) -> Tuple[float, float]:
    """(P(bit = 0), P(bit = 1)), unnormalised, of logical index bit `logical_bit` (qubit axis i of an N-qubit state
    is bit N-1-i). A local bit is a marginal reduction of every shard; a rank bit splits the ranks: each rank
    contributes its whole squared norm to the side its rank bit selects."""
    pos = phys_of[logical_bit]
    if pos < nl:
        if local is None:
            from . import engine
            p0, p1 = (float(v) for v in engine.marginal(shard, pos).cpu())
        else:
            p0, p1 = (float(v) for v in local(shard, pos))
    else:
        if local_norm2 is None:
            from . import engine
            n2 = float(engine.norm2(shard))
        else:
            n2 = float(local_norm2(shard))
        p0, p1 = (0.0, n2) if (rank >> (pos - nl)) & 1 else (n2, 0.0)
    total = _all_sum([p0, p1], shard, group)
    return total[0], total[1]


def expectation_diag(shard: torch.Tensor, local_diag: torch.Tensor, group=None, local: Callable = None) -> float:
    """sum_i d_i |psi_i|^2 for a real diagonal observable; `local_diag` is this rank's slice of the diagonal in
    the PHYSICAL layout of the shard (physical_diagonal)."""
    if local is None:
        from . import engine
        mine = float(engine.expectation_diag(shard, local_diag))
    else:
        mine = float(local(shard, local_diag))
    return _all_sum([mine], shard, group)[0]


def physical_diagonal(diag_logical: np.ndarray, phys_of: Sequence[int], nl: int, rank: int) -> np.ndarray:
    """This rank's slice of a diagonal given in logical flat order (index bit b = logical bit b), laid out like
    the shard: local physical index -> logical index through phys_of. For test-sized states (the whole diagonal
    is materialised); production observables are built per shard the same way."""
    nbits = len(phys_of)
    flat = np.asarray(diag_logical).reshape(-1)
    local = np.arange(1 << nl, dtype=np.int64) | (np.int64(rank) << nl)
    logical = np.zeros_like(local)
    for b, pos in enumerate(phys_of):
        logical |= ((local >> pos) & 1) << b
    assert flat.size == 1 << nbits
    return flat[logical]


def sample_indices(shard: torch.Tensor, uniforms: Sequence[float], phys_of: Sequence[int], nl: int, rank: int,
                   world: int, group=None, local_norm2: Callable = None, local_search: Callable = None) -> np.ndarray:
    """Basis-state indices (LOGICAL flat indices) drawn from |psi|^2 by inverting the cumulative distribution in
    physical order: the ranks' squared norms are gathered (P doubles), each uniform picks the rank whose span of
    the cumulative sum it falls into, that rank searches its own shard (engine.sample_search) and the indices are
    combined with one all-reduce. `uniforms` in [0, 1) must be the same on every rank (shared seed)."""
    if local_norm2 is None:
        from . import engine
        mine = float(engine.norm2(shard))
    else:
        mine = float(local_norm2(shard))
    totals = torch.zeros(world, dtype=torch.float64, device=shard.device)
    totals[rank] = mine
    dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)
    totals = totals.cpu().numpy()
    edges = np.concatenate([[0.0], np.cumsum(totals)])
    u = np.asarray(uniforms, dtype=np.float64) * edges[-1]
    owner = np.minimum(np.searchsorted(edges, u, side='right') - 1, world - 1)
    out = torch.zeros(len(u), dtype=torch.int64, device=shard.device)
    sel = np.nonzero(owner == rank)[0]
    if sel.size:
        # position inside this rank's span, rescaled to [0, 1) of the shard's own total
        inner = np.clip((u[sel] - edges[rank]) / totals[rank], 0.0, np.nextafter(1.0, 0.0))
        if local_search is None:
            from . import engine
            local_idx = np.asarray(engine.sample_search(engine.probabilities(shard).reshape(-1), inner), dtype=np.int64)
        else:
            local_idx = np.asarray(local_search(shard, inner), dtype=np.int64)
        physical = local_idx | (np.int64(rank) << nl)
        logical = np.zeros_like(physical)
        for b, pos in enumerate(phys_of):
            logical |= ((physical >> pos) & 1) << b
        out[torch.as_tensor(sel, device=out.device)] = torch.as_tensor(logical, device=out.device)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out.cpu().numpy()
