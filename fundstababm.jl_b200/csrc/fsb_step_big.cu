// fsb_step_big.cu -- the FSB_BIG build of the step kernels (fsb_step.cuh): shapes whose per-asset arrays do
// not fit in shared memory (BASELINE.json configs[3]: 2000 funds x 5000 assets).  Same source, same
// arithmetic; the kernels are named fsb_big_*_kernel and launched by fsb_abi.cu when DevState::big is set.
#define FSB_BIG 1
// wider event CTAs: the per-asset loops are 20x longer than on the default shape, and only 2 CTAs fit per SM
#define FSB_STEP_THREADS 256
#define FSB_E2_THREADS 256
#ifndef FSB_EV_CTAS
#define FSB_EV_CTAS 2  // (the event kernels' 107 KB of shared memory at the config-4 shape allow two CTAs per SM)
#endif
#ifdef FSB_BIG_TIMING  // experiments: per-phase cycles of the big kernels (scripts/phase_cycles.py --config4)
#define FSB_PHASE_TIMING 1
#define g_phase_cycles g_phase_cycles_big
#else
#undef FSB_PHASE_TIMING
#endif
#include "fsb_step.cuh"
#ifdef FSB_BIG_TIMING
extern "C" int fsb_debug_big_phase_cycles(unsigned long long *out, int reset)
{
    if (cudaMemcpyFromSymbol(out, fsb::g_phase_cycles_big, 32 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[32] = {}; if (cudaMemcpyToSymbol(fsb::g_phase_cycles_big, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
#endif
