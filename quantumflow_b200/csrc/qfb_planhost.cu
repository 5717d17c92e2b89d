// qfb_planhost.cu -- host-side support of the sweep planner (quantumflow_b200/planner.py). No device code.
//
// The planner chooses the tile of a sweep (which 2^M-amplitude slices a CTA holds on chip) by local search:
// exchange one tile bit for one outside bit, keep the exchange that lets the sweep execute more operators
// (Planner._refine_tile). Scoring a candidate tile is a scan over the next operators in program order with a
// handful of bit-mask tests per operator; the search scores a few thousand candidates per sweep, which is what
// this file runs natively (the Python statement of the same scan is tests/plan_emulator.py::count_executed,
// and a test checks that both agree).
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

}  // extern "C"
