// qfb_planhost.cu -- host-side support of the sweep planner (quantumflow_b200/planner.py). No device code.
//
// The planner chooses the tile of a sweep (which 2^M-amplitude slices a CTA holds on chip) by local search:
// exchange one tile bit for one outside bit, keep the exchange that lets the sweep execute more operators
// (Planner._refine_tile). Scoring a candidate tile is a scan over the next operators in program order with a
// handful of bit-mask tests per operator; the search scores a few thousand candidates per sweep, which is what
// this file runs natively (the Python statement of the same scan is tests/plan_emulator.py::count_executed,
// and a test checks that both agree).
#include <stdlib.h>
#include "qfb_common.cuh"

namespace qfb {

struct PlanOp {
    uint64_t mix;     // index bits the operator mixes
    uint64_t diag;    // index bits it reads as controls / phase-term bits
    double cost;      // planner work units
    uint32_t bytes;   // upper bound of its records in the sweep
};

// number of operators that touch a bit and that a sweep over `tmask` executes: an operator joins when it mixes
// tile bits only, commutes with everything deferred before it (da: bits touched, dm: bits mixed by the deferred
// operators) and fits the work and size caps of a sweep
static int count_executed(const uint64_t *mix, const uint64_t *diag, const double *cost, const uint32_t *bytes,
                          int nops, uint64_t tmask, uint64_t fmask, double max_cost, int64_t room) {
    const uint64_t allow = tmask & ~fmask;
    uint64_t da = 0, dm = 0;
    double work = 0.0;
    int count = 0;
    for (int i = 0; i < nops; ++i) {
        const uint64_t mm = mix[i], dd = diag[i];
        if ((mm & da) || (dd & dm) || (mm & ~allow) || (count && work + cost[i] > max_cost)) {
            da |= mm | dd;
            dm |= mm;
            if (!(allow & ~da)) break;   // no tile bit is open for mixing any more
            continue;
        }
        room -= bytes[i];
        if (room < 0) break;
        work += cost[i];
        if (mm | dd) ++count;
    }
    return count;
}

// ---- two-sweep look-ahead: what does a tile leave for the NEXT sweep? ----
struct PlanCtx {
    const uint64_t *mix, *diag;
    const double *cost;
    const uint32_t *bytes;
    int nbits, low_bits, tile_bits;
    uint64_t fmask;
    double max_cost;
    int64_t room;
    int passes;
};

// same admission rule as count_executed over an index list; the deferred indices are written to `rest`
static int closure_indexed(const PlanCtx &c, const int *idx, int n, uint64_t tmask, int *rest, int *nrest) {
    const uint64_t allow = tmask & ~c.fmask;
    uint64_t da = 0, dm = 0;
    double work = 0.0;
    int64_t room = c.room;
    int count = 0, nr = 0;
    bool started = false;
    for (int k = 0; k < n; ++k) {
        const int i = idx[k];
        const uint64_t mm = c.mix[i], dd = c.diag[i];
        if ((mm & da) || (dd & dm) || (mm & ~allow) || (started && work + c.cost[i] > c.max_cost)) {
            rest[nr++] = i;
            da |= mm | dd;
            dm |= mm;
            continue;
        }
        if (room - (int64_t)c.bytes[i] < 0) {
            for (; k < n; ++k) rest[nr++] = idx[k];
            break;
        }
        room -= c.bytes[i];
        started = true;
        work += c.cost[i];
        if (mm | dd) ++count;
    }
    *nrest = nr;
    return count;
}

static int count_indexed(const PlanCtx &c, const int *idx, int n, uint64_t tmask) {
    const uint64_t allow = tmask & ~c.fmask;
    uint64_t da = 0, dm = 0;
    double work = 0.0;
    int64_t room = c.room;
    int count = 0;
    for (int k = 0; k < n; ++k) {
        const int i = idx[k];
        const uint64_t mm = c.mix[i], dd = c.diag[i];
        if ((mm & da) || (dd & dm) || (mm & ~allow) || (count && work + c.cost[i] > c.max_cost)) {
            da |= mm | dd;
            dm |= mm;
            if (!(allow & ~da)) break;
            continue;
        }
        room -= c.bytes[i];
        if (room < 0) break;
        work += c.cost[i];
        if (mm | dd) ++count;
    }
    return count;
}

// the greedy walk of Planner._form_sweep (no randomisation): the first operators that fit claim the tile
static uint64_t greedy_tile(const PlanCtx &c, const int *idx, int n) {
    uint64_t tile = (1ull << c.low_bits) - 1ull, da = 0, dm = 0;
    double work = 0.0;
    int64_t room = c.room;
    bool started = false;
    for (int k = 0; k < n; ++k) {
        const int i = idx[k];
        const uint64_t mm = c.mix[i], dd = c.diag[i];
        bool ok = !((mm & da) || (dd & dm)) && !(mm & c.fmask);
        if (ok && mm && __builtin_popcountll(tile | mm) > c.tile_bits) ok = false;
        if (ok && started && work + c.cost[i] > c.max_cost) ok = false;
        if (ok && room - (int64_t)c.bytes[i] < 0) break;
        if (ok) {
            started = true;
            work += c.cost[i];
            room -= c.bytes[i];
            tile |= mm;
        } else {
            da |= mm | dd;
            dm |= mm;
        }
    }
    for (int b = 0; b < c.nbits && __builtin_popcountll(tile) < c.tile_bits; ++b)
        if (!((tile >> b) & 1ull) && !((c.fmask >> b) & 1ull)) tile |= 1ull << b;
    return tile;
}

static int refine_indexed(const PlanCtx &c, const int *idx, int n, uint64_t &tmask, uint64_t keep) {
    int best = count_indexed(c, idx, n, tmask);
    for (int pass = 0; pass < c.passes; ++pass) {
        const uint64_t base = tmask;
        bool improved = false;
        for (int bi = 0; bi < c.nbits; ++bi) {
            if (!((base >> bi) & 1ull) || ((keep >> bi) & 1ull)) continue;
            const uint64_t without = base & ~(1ull << bi);
            for (int bo = 0; bo < c.nbits; ++bo) {
                if (((base >> bo) & 1ull) || ((c.fmask >> bo) & 1ull)) continue;
                const int m = count_indexed(c, idx, n, without | (1ull << bo));
                if (m > best) {
                    best = m;
                    tmask = without | (1ull << bo);
                    improved = true;
                }
            }
        }
        if (!improved) break;
    }
    return best;
}

// operators this sweep executes + operators the best next sweep (greedy tile + local search) then executes
static int two_sweep_score(const PlanCtx &c, const int *idx, int n, uint64_t tmask, int *rest) {
    int nrest = 0;
    const int now = closure_indexed(c, idx, n, tmask, rest, &nrest);
    if (nrest == 0) return now + (1 << 20);
    uint64_t next = greedy_tile(c, rest, nrest);
    return now + refine_indexed(c, rest, nrest, next, (1ull << c.low_bits) - 1ull);
}

}  // namespace qfb

using namespace qfb;

extern "C" {

int qfb_plan_count_executed(const uint64_t *mix, const uint64_t *diag, const double *cost, const uint32_t *bytes,
                            int nops, uint64_t tmask, uint64_t fmask, double max_cost, int64_t room, int *count_out) {
    QFB_CHECK_ARG(nops >= 0 && (nops == 0 || (mix && diag && cost && bytes)) && count_out,
                  "qfb_plan_count_executed: bad arguments");
    *count_out = count_executed(mix, diag, cost, bytes, nops, tmask, fmask, max_cost, room);
    return QFB_OK;
}

int qfb_plan_refine_tile(const uint64_t *mix, const uint64_t *diag, const double *cost, const uint32_t *bytes,
                         int nops, int nbits, uint64_t tmask, uint64_t fmask, uint64_t keep, double max_cost,
                         int64_t room, int passes, uint64_t *tmask_out, int *count_out) {
    QFB_CHECK_ARG(nops >= 0 && (nops == 0 || (mix && diag && cost && bytes)) && tmask_out,
                  "qfb_plan_refine_tile: bad arguments");
    QFB_CHECK_ARG(nbits >= 1 && nbits <= 62, "qfb_plan_refine_tile: nbits=%d out of range", nbits);
    int best = count_executed(mix, diag, cost, bytes, nops, tmask, fmask, max_cost, room);
    for (int pass = 0; pass < passes; ++pass) {
        const uint64_t base = tmask;
        bool improved = false;
        for (int bi = 0; bi < nbits; ++bi) {
            if (!((base >> bi) & 1ull) || ((keep >> bi) & 1ull)) continue;
            const uint64_t without = base & ~(1ull << bi);
            for (int bo = 0; bo < nbits; ++bo) {
                if (((base >> bo) & 1ull) || ((fmask >> bo) & 1ull)) continue;
                const uint64_t cand = without | (1ull << bo);
                const int n = count_executed(mix, diag, cost, bytes, nops, cand, fmask, max_cost, room);
                if (n > best) {
                    best = n;
                    tmask = cand;
                    improved = true;
                }
            }
        }
        if (!improved) break;
    }
    *tmask_out = tmask;
    if (count_out) *count_out = best;
    return QFB_OK;
}

int qfb_plan_refine_tile_lookahead(const uint64_t *mix, const uint64_t *diag, const double *cost,
                                   const uint32_t *bytes, int nops, int nbits, int low_bits, int tile_bits,
                                   uint64_t tmask, uint64_t fmask, uint64_t keep, double max_cost, int64_t room,
                                   int passes, int lookahead_passes, uint64_t *tmask_out, int *score_out) {
    QFB_CHECK_ARG(nops >= 0 && (nops == 0 || (mix && diag && cost && bytes)) && tmask_out,
                  "qfb_plan_refine_tile_lookahead: bad arguments");
    QFB_CHECK_ARG(nbits >= 1 && nbits <= 62 && low_bits >= 0 && low_bits <= tile_bits && tile_bits <= nbits,
                  "qfb_plan_refine_tile_lookahead: nbits=%d low_bits=%d tile_bits=%d out of range", nbits, low_bits,
                  tile_bits);
    PlanCtx c{mix, diag, cost, bytes, nbits, low_bits, tile_bits, fmask, max_cost, room, passes};
    int *idx = (int *)malloc(sizeof(int) * (size_t)(2 * nops + 2));
    QFB_CHECK_ARG(idx, "qfb_plan_refine_tile_lookahead: out of memory");
    int *rest = idx + nops + 1;
    for (int i = 0; i < nops; ++i) idx[i] = i;
    int best = two_sweep_score(c, idx, nops, tmask, rest);
    for (int pass = 0; pass < lookahead_passes; ++pass) {
        const uint64_t base = tmask;
        bool improved = false;
        for (int bi = 0; bi < nbits; ++bi) {
            if (!((base >> bi) & 1ull) || ((keep >> bi) & 1ull)) continue;
            const uint64_t without = base & ~(1ull << bi);
            for (int bo = 0; bo < nbits; ++bo) {
                if (((base >> bo) & 1ull) || ((fmask >> bo) & 1ull)) continue;
                const int sc = two_sweep_score(c, idx, nops, without | (1ull << bo), rest);
                if (sc > best) {
                    best = sc;
                    tmask = without | (1ull << bo);
                    improved = true;
                }
            }
        }
        if (!improved) break;
    }
    free(idx);
    *tmask_out = tmask;
    if (score_out) *score_out = best;
    return QFB_OK;
}

}  // extern "C"
