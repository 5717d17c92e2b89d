"""State vectors sharded over P = 2^p GPUs of one box (one process per GPU, torch.distributed).

The reference has no distributed path; this is the B200-native scaling axis. Flat index =
(rank bits | local bits): rank r owns the contiguous slice [r 2^nl, (r+1) 2^nl). A logical->physical bit map is
kept on the host, so "swapping back" is never needed.

* diagonal operators and control qubits on rank bits cost no communication: the kernels resolve them from
  `index_hi` (= the rank), see qfb_apply_diag / qfb_run_plan;
* an operator that MIXES a rank bit triggers a remap: k rank bits are exchanged with the k top local bits by a
  pairwise block exchange (for k = p this is the all-to-all of SURVEY 8e) over NCCL / NVLink. The outgoing logical
  qubits are first moved to the top local positions by an in-place bit permutation that rides on the final store
  of the stage's last sweep (planner.attach_permutation: no extra pass over the shard);
* the exchange is IN PLACE: block j of the shard is swapped with the matching block of peer (mine ^ s) in step s
  (XOR pairing, so every pair agrees on the order), chunk by chunk through two small staging buffers - chunk
  i+1 is on the wire while chunk i is copied out of its staging buffer. Memory: shard + 2 staging chunks, which
  is what lets a 2^33-amplitude (128 GiB) shard live on a 180 GB GPU;
* which qubits become global is decided Belady-style: the ones whose next mixing use is farthest away.

`ShardedCircuit` only needs three callables for the local work (run segments, permute bits, allocate scratch),
so the scheduling and exchange logic is testable on CPU with gloo and the oracle standing in for the kernels
(tests/test_sharded_cpu.py); on GPU the callables are the libqfb200 paths.
"""
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
    """Local work between two remaps: operators with PHYSICAL bit positions, planned for nl local bits, followed
    by the in-place local bit permutation `final_perm` (dst local bit j <- src local bit final_perm[j]; None =
    identity) that prepares the next remap."""
    __slots__ = ('bitops', 'segments', 'final_perm')

    def __init__(self, bitops: List[BitOp], final_perm: Optional[List[int]] = None):
        self.bitops = bitops
        self.segments = None
        self.final_perm = final_perm


class Remap:
    """Exchange rank bits `rank_positions` (positions inside the rank index) with the top-k local bits."""
    __slots__ = ('rank_positions',)

    def __init__(self, rank_positions: List[int]):
        self.rank_positions = rank_positions


def schedule(nbits: int, p: int, bitops: Sequence[BitOp]) -> Tuple[List[object], List[int]]:
    """Split `bitops` (logical bit positions, program order) into Stage / Remap steps.

    Returns (steps, final phys_of) where phys_of[logical bit] = physical bit; physical bits >= nbits - p are rank
    bits. Deterministic and identical on every rank."""
    nl = nbits - p
    phys_of = list(range(nbits))                       # identity: the top p qubits' bits are global
    info = [(_mixing_and_diag_bits(np.asarray(m), list(b))) for m, b in bitops]
    remaining = list(range(len(bitops)))
    steps: List[object] = []
    while remaining:
        # greedy stage: everything executable under the current map, reordered only across commuting operators
        stage_ops: List[BitOp] = []
        deferred: List[int] = []
        def_any: set = set()
        def_mix: set = set()
        for i in remaining:
            mix, diag = info[i]
            conflict = bool(mix & def_any) or bool(diag & def_mix)
            local = all(phys_of[b] < nl for b in mix)
            if not conflict and local:
                mat, bits = bitops[i]
                stage_ops.append((mat, [phys_of[b] for b in bits]))
            else:
                deferred.append(i)
                def_any |= mix | diag
                def_mix |= mix
        if stage_ops:
            steps.append(Stage(stage_ops))
        remaining = deferred
        if not remaining:
            break
        if p == 0:
            raise RuntimeError('unschedulable operator')
        # Belady: keep local the bits that are mixed soonest; the p bits used farthest in the future go global
        next_use: Dict[int, int] = {}
        for order, i in enumerate(remaining):
            for b in info[i][0]:
                next_use.setdefault(b, order)
        far = sorted(range(nbits), key=lambda b: (-next_use.get(b, 1 << 60), -phys_of[b]))
        new_global = set(far[:p])
        cur_global = {b for b in range(nbits) if phys_of[b] >= nl}
        outgoing = sorted(new_global - cur_global, key=lambda b: phys_of[b])     # local -> rank
        incoming = sorted(cur_global - new_global, key=lambda b: phys_of[b])     # rank -> local
        k = len(outgoing)
        if k == 0:
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
            new_logical_at = {j: logical_at[perm[j]] for j in range(nl)}
            for j, b in new_logical_at.items():
                phys_of[b] = j
        # exchange: rank bit position t (physical nl + t) <-> local position top[i], pairing in ascending order
        rank_positions = sorted(phys_of[b] - nl for b in incoming)
        logical_at = {phys_of[b]: b for b in range(nbits)}
        for i, t in enumerate(rank_positions):
            b_local, b_rank = logical_at[top[i]], logical_at[nl + t]
            phys_of[b_local], phys_of[b_rank] = nl + t, top[i]
        if local_perm is not None:
            if not steps or not isinstance(steps[-1], Stage):
                steps.append(Stage([]))
            steps[-1].final_perm = local_perm
        steps.append(Remap(rank_positions))
    return steps, phys_of


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
        self.steps, self.final_phys_of = schedule(nqubits, p, bitops)
        self._plan_args = dict(tile_bits=tile_bits, low_bits=low_bits, max_cost=max_cost)
        self._run_stage = run_stage or self._run_stage_gpu
        self._permute = permute            # test double only (out of place); the GPU path fuses it into the plan
        self._staging_bytes = int(staging_bytes)
        self._staging = None
        self._comm_seconds = 0.0
        self._comm_bytes = 0
        self._remaps = 0
        self._executions = 0
        for st in self.steps:
            if isinstance(st, Stage):
                st.segments = planner.build_segments(self.nl, st.bitops, final_perm=st.final_perm,
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

    def comm_ms_per_step(self) -> float:
        return 1e3 * self._comm_seconds / max(1, self._executions)

    def comm_summary(self) -> dict:
        ex = max(1, self._executions)
        return {'remaps_per_step': self._remaps / ex, 'bytes_sent_per_rank_per_step': self._comm_bytes / ex,
                'ms_per_step': self.comm_ms_per_step(),
                'note': 'in-place pairwise block exchange (isend/irecv, XOR pairing) of k rank bits with the top-k '
                        'local bits, chunked through two staging buffers; the local bit permutation is fused '
                        'into the last sweep of the preceding stage; not overlapped with compute yet'}

    def reset_comm_counters(self) -> None:
        self._comm_seconds, self._comm_bytes, self._remaps, self._executions = 0.0, 0, 0, 0

    # ---- execution ---------------------------------------------------------------------------------
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

    def execute(self, shard: torch.Tensor) -> torch.Tensor:
        """Runs the circuit in place on this rank's shard and returns it (the same tensor)."""
        assert shard.numel() == 1 << self.nl and shard.is_contiguous()
        timed = shard.is_cuda
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
            self._exchange(shard, st.rank_positions)
            self._remaps += 1
            if timed:
                ev1.record()
                ev1.synchronize()
                self._comm_seconds += ev0.elapsed_time(ev1) * 1e-3
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
