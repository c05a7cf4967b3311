// fsb_step_big.cu -- the FSB_BIG build of the step kernels (fsb_step.cuh): shapes whose per-asset arrays do
// not fit in shared memory (BASELINE.json configs[3]: 2000 funds x 5000 assets).  Same source, same
// arithmetic; the kernels are named fsb_big_*_kernel and launched by fsb_abi.cu when DevState::big is set.
#define FSB_BIG 1
// wider event CTAs: the per-asset loops are 20x longer than on the default shape, and only 2 CTAs fit per SM
#define FSB_STEP_THREADS 256
#define FSB_E2_THREADS 256
#define FSB_EV_CTAS 2
#undef FSB_PHASE_TIMING
#include "fsb_step.cuh"
