# time_reference.jl -- the reference's own CPU time for BASELINE.json configs[0] (SURVEY.md 8d-1):
# one warm-up call, then one timed `runmodel()` of the default Params, reported as fund-steps/s next
# to bench.py's numbers.  The reference has no threaded code; BLAS threads are printed.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI (no Julia in the build image); bench.py --impl reference times
# the C restatement (oracle/) instead and says so.
#
#   JULIA_NUM_THREADS=1 julia --project=<FundStabABM.jl checkout> time_reference.jl
using LinearAlgebra
using FundStabABM
p = FundStabABM.Params.default()
steps = p.bigt - (p.perfwindow[end] + 1)
FundStabABM.runmodel(p)                      # warm-up (compilation)
t = @elapsed FundStabABM.runmodel(p)
println("julia threads: ", Threads.nthreads(), "  BLAS threads: ", BLAS.get_num_threads())
println("fund_steps_per_sec: ", p.bigk * steps / t, "  (", steps, " steps, ", p.bigk, " funds, ", t, " s)")
