// fsb_step_fixed.cu -- the FSB_FIXED build of the step kernels (fsb_step.cuh): exactly the default Params shape
// (src/params.jl:5-8,24: 250 funds x 250 stocks x 1000 investors; reviewprobability 1/63 -> event capacity 80), with the
// shapes and every shared-memory offset of the step context as compile-time constants.  Same source, same arithmetic;
// the kernels are named fsb_fix_*_kernel and fsb_abi.cu launches the event and choice kernels from here when an engine
// has exactly this shape (anything else -- other shapes, sweeps with another review probability, FSB_EVENT_ROWS -- runs
// the generic build).
#define FSB_FIXED 1
#undef FSB_PHASE_TIMING
#include "fsb_step.cuh"
