// qfb_jit.h -- sweep-specialised kernels (qfb_jit.cu): internal interface used by the plan executor (qfb_sweep.cu).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include <cuda_runtime.h>

namespace qfb {

struct JitSource {
    std::string ptx;              // depends on the STRUCTURE of the sweep only (cache key of the compiled image)
    std::vector<double> coef;     // the sweep's coefficients in the order the code reads them (module global qfb_coef)
    size_t coef_bytes = 0;
    int threads = 0;
    size_t smem_bytes = 0;
    int nholes = 0;
    int groups = 1;               // tiles a CTA works on side by side
};

struct JitSweep;                  // a loaded module + launch geometry

// PTX text and coefficients of the sweep record at `rec` (a record of a VALIDATED plan)
// (fix_mask: index bits outside the tile that a launch fixes through the kernel's p_fix parameter, see jit_build_variant)
int jit_generate(const uint8_t *rec, int nbits, int tile_bits, int reg_bits, JitSource &out, std::string &err,
                 uint64_t fix_mask = 0);
// PTX -> sm_100a image with the statically linked PTX compiler (no GPU, no driver needed); `log` = ptxas -v output
int jit_compile(const std::string &ptx, std::vector<char> &cubin, std::string &log);
// all sweeps of a plan: generate + compile in parallel (images cached per process by PTX text), load into the current
// context, write the coefficient banks
int jit_build_plan(const uint8_t *plan, const std::vector<size_t> &offsets, int nbits, int tile_bits, int reg_bits,
                   std::vector<JitSweep *> &out);
// one sweep for launches over a slice of the state: the index bits `fix_mask` come from jit_launch's fix_value
int jit_build_variant(const uint8_t *rec, int nbits, int tile_bits, int reg_bits, uint64_t fix_mask, JitSweep **out);
int jit_launch(JitSweep *s, void *state, uint64_t hi_shifted, cudaStream_t st, uint64_t fix_value = 0,
               int ctas_per_sm = 0);
void jit_destroy(JitSweep *s);
void jit_cache_stats(uint64_t *hits, uint64_t *misses);

}  // namespace qfb
