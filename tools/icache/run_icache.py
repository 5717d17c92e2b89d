#!/usr/bin/env python
"""Instruction-cache probe (GPU part): runs icache_<KB>.cubin with 3 CTAs of 128 threads per SM and prints the FP64
issue rate per SM against the size of the loop body. Usage: python tools/icache/run_icache.py"""
import ctypes, glob, json, os, re
import numpy as np
import torch
from cuda import cuda

HERE = os.path.dirname(os.path.abspath(__file__))
torch.cuda.init(); torch.zeros(1, device='cuda')
def chk(r):
    if r[0] != cuda.CUresult.CUDA_SUCCESS: raise RuntimeError(str(r[0]))
    return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)
buf = torch.zeros(1 << 22, dtype=torch.float64, device='cuda')
sms = torch.cuda.get_device_properties(0).multi_processor_count
files = sorted(glob.glob(os.path.join(HERE, 'icache_*.cubin')), key=lambda f: int(re.findall(r'(\d+)\.cubin', f)[0]))
for ctas_per_sm in (3, 1):
    for f in files:
        kb = int(re.findall(r'(\d+)\.cubin', f)[0])
        mod = chk(cuda.cuModuleLoadData(open(f, 'rb').read()))
        fn = chk(cuda.cuModuleGetFunction(mod, b'k'))
        iters = max(4, 4096 // kb)
        params = ((buf.data_ptr(), iters, 1e-9, 1e-9), (ctypes.c_void_p, ctypes.c_uint32, ctypes.c_double, ctypes.c_double))
        grid = sms * ctas_per_sm
        def launch():
            chk(cuda.cuLaunchKernel(fn, grid, 1, 1, 128, 1, 1, 0, 0, params, 0))
        launch(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        dfma = grid * 4 * iters * kb * 64
        print(json.dumps({'loop_body_KiB': kb, 'ctas_per_sm': ctas_per_sm, 'ms': round(ms, 3),
                          'dfma_warp_instr_per_clk_per_sm_at_1965MHz': round(dfma / (ms * 1e-3) / sms / 1.965e9, 3)}), flush=True)
        chk(cuda.cuModuleUnload(mod))
