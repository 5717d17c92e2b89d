// qfb_planhost.cu -- host-side support of the sweep planner (quantumflow_b200/planner.py). No device code.
//
// The planner chooses the tile of a sweep (which 2^M-amplitude slices a CTA holds on chip) by local search:
// exchange one tile bit for one outside bit, keep the exchange that lets the sweep execute more operators
// (Planner._refine_tile). Scoring a candidate tile is a scan over the next operators in program order with a
// handful of bit-mask tests per operator; the search scores a few thousand candidates per sweep, which is what
// this file runs natively (the Python statement of the same scan is tests/plan_emulator.py::count_executed,
// and a test checks that both agree).
#include <math.h>
#include <stdlib.h>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>
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

// ---- candidates of one search pass side by side --------------------------------------------------------------
// A pass of the tile search scores ~150-250 candidate tiles independently; the scores go into an array and the
// winner is picked afterwards in the serial order, so the result does not depend on the number of threads.
// The helper threads live for ONE call of a search function (nothing persists in the process: safe across fork,
// nothing to shut down) and wait for the next pass of that call by polling a counter (a call lasts a few
// milliseconds). QFB_PLAN_THREADS / qfb_plan_set_threads (default: up to 4, less when the ranks of a multi-GPU job
// share the cores) = 1 keeps everything on the calling thread.
static std::atomic<int> g_plan_threads{0};      // 0 = not decided yet

static int plan_threads() {
    int cached = g_plan_threads.load(std::memory_order_relaxed);
    if (cached) return cached;
    int want = 4;
    const char *v = getenv("QFB_PLAN_THREADS");
    if (v && *v) {
        want = atoi(v);
    } else {
        // one process per GPU (torchrun): every rank plans at the same time, share the cores
        const char *w = getenv("LOCAL_WORLD_SIZE");
        const int ranks = (w && *w) ? atoi(w) : 1;
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && ranks > 0 && want > hw / ranks) want = hw / ranks;
    }
    if (want < 1) want = 1;
    if (want > 16) want = 16;
    g_plan_threads.store(want, std::memory_order_relaxed);
    return want;
}

class PassWorkers {
public:
    // weight = rough cost of one candidate (operators scanned): helper threads are only started for passes that are
    // worth a few hundred microseconds (small states have few candidates and short operator lists)
    PassWorkers(int nthreads, int64_t weight) : want_(nthreads), weight_(weight) {}
    ~PassWorkers() {
        if (workers_.empty()) return;
        quit_.store(true, std::memory_order_relaxed);
        generation_.fetch_add(1, std::memory_order_release);
        for (auto &w : workers_) w.join();
    }
    int nthreads() const { return want_; }
    // body(i, worker) for every i in [0, n): worker w takes i = w, w + T, ... (the caller is worker 0)
    void run(int n, const std::function<void(int, int)> &body) {
        if (want_ > 1 && workers_.empty() && n >= 2 * want_ && (int64_t)n * weight_ >= 40000) start();
        const int T = 1 + (int)workers_.size();
        if (T == 1) {
            for (int i = 0; i < n; ++i) body(i, 0);
            return;
        }
        body_ = &body;
        n_ = n;
        done_.store(0, std::memory_order_relaxed);
        generation_.fetch_add(1, std::memory_order_release);
        for (int i = 0; i < n; i += T) body(i, 0);
        while (done_.load(std::memory_order_acquire) != T - 1) std::this_thread::yield();
    }

private:
    void start() {
        for (int t = 1; t < want_; ++t) {
            try {
                workers_.emplace_back([this, t]() { loop(t); });
            } catch (...) {
                break;                // no more threads to be had: fewer workers
            }
        }
        stride_ = 1 + (int)workers_.size();
    }
    void loop(int t) {
        int seen = 0;
        for (;;) {
            int g;
            while ((g = generation_.load(std::memory_order_acquire)) == seen) std::this_thread::yield();
            seen = g;
            if (quit_.load(std::memory_order_relaxed)) return;
            // stride_ is final here: run() publishes a generation only after start() has returned
            for (int i = t; i < n_; i += stride_) (*body_)(i, t);
            done_.fetch_add(1, std::memory_order_release);
        }
    }
    int want_;
    int64_t weight_;
    int stride_ = 1;
    int n_ = 0;
    const std::function<void(int, int)> *body_ = nullptr;
    std::vector<std::thread> workers_;
    std::atomic<int> generation_{0};
    std::atomic<int> done_{0};
    std::atomic<bool> quit_{false};
};

// the exchanges of one pass around `base`: (tile bit outside `keep`) x (outside bit not in fmask), in the order the
// serial search visits them
static void pass_candidates(uint64_t base, uint64_t keep, uint64_t fmask, int nbits, std::vector<uint64_t> &out) {
    out.clear();
    for (int bi = 0; bi < nbits; ++bi) {
        if (!((base >> bi) & 1ull) || ((keep >> bi) & 1ull)) continue;
        const uint64_t without = base & ~(1ull << bi);
        for (int bo = 0; bo < nbits; ++bo) {
            if (((base >> bo) & 1ull) || ((fmask >> bo) & 1ull)) continue;
            out.push_back(without | (1ull << bo));
        }
    }
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

int qfb_plan_set_threads(int nthreads) {
    g_plan_threads.store(nthreads > 16 ? 16 : (nthreads > 0 ? nthreads : 0), std::memory_order_relaxed);
    return QFB_OK;
}

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
    std::vector<uint64_t> cands;
    std::vector<int> score;
    PassWorkers pool(plan_threads(), nops);
    const std::function<void(int, int)> body = [&](int i, int) {
        score[(size_t)i] = count_executed(mix, diag, cost, bytes, nops, cands[(size_t)i], fmask, max_cost, room);
    };
    for (int pass = 0; pass < passes; ++pass) {
        pass_candidates(tmask, keep, fmask, nbits, cands);
        score.assign(cands.size(), 0);
        pool.run((int)cands.size(), body);
        bool improved = false;
        for (size_t i = 0; i < cands.size(); ++i) {
            if (score[i] > best) {
                best = score[i];
                tmask = cands[i];
                improved = true;
            }
        }
        if (!improved) break;
    }
    *tmask_out = tmask;
    if (count_out) *count_out = best;
    return QFB_OK;
}

// The split of a sweep's operators into rounds of `reg_bits` register bits (Planner._split_rounds states the same
// greedy in Python; tests/test_planner.py compares the two): operators in program order (reversed with `backward`:
// latest-possible rounds), an operator joins a round when it commutes with everything deferred before it and, if it
// mixes bits, when their tile positions (posmask; 0xffffffff marks a phase term) still fit the round's register bits;
// the first round keeps the low tile positions on the lanes. `rnd` (nrnd numbers in [0, 1), the stream of the planner's
// random.Random(trial)) makes a randomised variant: an operator that needs a NEW register bit is admitted with
// probability p_new. Outputs: rounds, the cost of the last round's operators (summed like Python's sum()), and on
// request the round of every operator and the register positions of every round. *consumed_out = -1: the random
// stream ran out (the caller falls back to its own loop).
int qfb_plan_split_rounds(const uint64_t *mix, const uint64_t *diag, const uint32_t *posmask, const double *cost,
                          int nops, int reg_bits, int low_bits, const double *rnd, int nrnd, double p_new,
                          int backward, int *round_of, uint32_t *regs_of_round, int max_rounds, int *nrounds_out,
                          double *tail_out, int *consumed_out) {
    QFB_CHECK_ARG(nops >= 0 && (nops == 0 || (mix && diag && posmask && cost)) && nrounds_out,
                  "qfb_plan_split_rounds: bad arguments");
    QFB_CHECK_ARG(reg_bits >= 1 && reg_bits <= 8 && low_bits >= 0 && low_bits < 32 && (rnd || nrnd == 0),
                  "qfb_plan_split_rounds: reg_bits=%d low_bits=%d out of range", reg_bits, low_bits);
    const uint32_t lowmask = (1u << low_bits) - 1u;
    std::vector<int> remaining((size_t)nops), deferred;
    std::vector<int> rof((size_t)nops, -1);
    std::vector<uint32_t> regs;          // register-bit positions of every round, processing orientation
    std::vector<int> nchosen;
    for (int k = 0; k < nops; ++k) remaining[(size_t)k] = backward ? nops - 1 - k : k;
    deferred.reserve((size_t)nops);
    int used = 0;
    bool first = true, starved = false;
    while (!remaining.empty()) {
        // cannot happen: every round after the first takes at least its first operator
        QFB_CHECK_ARG((int)regs.size() <= nops + 2, "qfb_plan_split_rounds: no progress");
        uint32_t rm = 0;
        int nregs = 0, count = 0;
        uint64_t da = 0, dm = 0;
        const int r = (int)regs.size();
        deferred.clear();
        for (int i : remaining) {
            const uint64_t mm = mix[i], dd = diag[i];
            bool ok = !((mm & da) || (dd & dm));
            if (ok && posmask[i] != 0xffffffffu) {          // a mixing operator (phase terms carry 0xffffffff)
                const uint32_t pm = posmask[i];
                const uint32_t need = pm & ~rm;
                const int cnt = __builtin_popcount(need);
                if (first && (pm & lowmask)) ok = false;
                else if (nregs + cnt > reg_bits) ok = false;
                else if (need && rm && rnd) {
                    if (used >= nrnd) { starved = true; ok = false; }
                    else if (rnd[used++] > p_new) ok = false;
                }
                if (ok) { rm |= need; nregs += cnt; }
            }
            if (ok) { rof[(size_t)i] = r; ++count; }
            else { deferred.push_back(i); da |= mm | dd; dm |= mm; }
        }
        if (starved) break;
        regs.push_back(rm);
        nchosen.push_back(count);
        remaining.swap(deferred);
        first = false;
    }
    if (consumed_out) *consumed_out = starved ? -1 : used;
    if (starved) { *nrounds_out = 0; return QFB_OK; }
    if (regs.empty()) { regs.push_back(0); nchosen.push_back(0); }
    if (regs.back() & lowmask) { regs.push_back(0); nchosen.push_back(0); }
    int shift = 0;
    if (regs.size() > 1 && nchosen[0] == 0 && !(regs[1] & lowmask)) shift = 1;
    const int nr = (int)regs.size() - shift;
    *nrounds_out = nr;
    // final orientation: a backward split is the mirror's rounds in reverse order
    // cost of the last round's operators, summed like CPython's sum() of floats (Neumaier's compensated sum)
    double tail = 0.0, comp = 0.0;
    bool any = false;
    for (int i = 0; i < nops; ++i) {
        int r = rof[(size_t)i] - shift;
        if (backward) r = nr - 1 - r;
        if (round_of) round_of[i] = r;
        if (r != nr - 1) continue;
        const double x = cost[i];
        if (!any) { tail = 0.0 + x; any = true; continue; }
        const double t = tail + x;
        if (fabs(tail) >= fabs(x)) comp += (tail - t) + x;
        else comp += (x - t) + tail;
        tail = t;
    }
    if (comp != 0.0 && isfinite(comp)) tail += comp;
    if (tail_out) *tail_out = tail;
    if (regs_of_round) {
        QFB_CHECK_ARG(nr <= max_rounds, "qfb_plan_split_rounds: %d rounds, room for %d", nr, max_rounds);
        for (int r = 0; r < nr; ++r) regs_of_round[backward ? nr - 1 - r : r] = regs[(size_t)(r + shift)];
    }
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
    std::vector<int> idx((size_t)nops + 1);
    for (int i = 0; i < nops; ++i) idx[(size_t)i] = i;
    PassWorkers pool(plan_threads(), (int64_t)nops * 64);
    std::vector<std::vector<int>> rest((size_t)pool.nthreads(), std::vector<int>((size_t)nops + 1));
    int best = two_sweep_score(c, idx.data(), nops, tmask, rest[0].data());
    std::vector<uint64_t> cands;
    std::vector<int> score;
    for (int pass = 0; pass < lookahead_passes; ++pass) {
        pass_candidates(tmask, keep, fmask, nbits, cands);
        score.assign(cands.size(), 0);
        pool.run((int)cands.size(), [&](int i, int worker) {
            score[(size_t)i] = two_sweep_score(c, idx.data(), nops, cands[(size_t)i], rest[(size_t)worker].data());
        });
        bool improved = false;
        for (size_t i = 0; i < cands.size(); ++i) {
            if (score[i] > best) {
                best = score[i];
                tmask = cands[i];
                improved = true;
            }
        }
        if (!improved) break;
    }
    *tmask_out = tmask;
    if (score_out) *score_out = best;
    return QFB_OK;
}

}  // extern "C"
