# FundStabB200.jl -- the `ccall` glue a FundStabABM.jl maintainer adds to route the hot path
# (Func.initialise, src/functions.jl:417; Func.modelrun, src/functions.jl:446) through
# libfsb_b200.so (include/fsb.h).  No model logic lives here: every function marshals arrays that
# Julia already stores column-major and turns a non-zero status into an error.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The same symbols are
# exercised from Python (fundstababm.jl_b200/_lib.py) by tests/.
module FundStabB200

const LIB = get(ENV, "FSB_LIB", "libfsb_b200.so")

# mirrors `struct fsb_params` of include/fsb.h (Params.default, src/params.jl:4-31)
struct FsbParams
    bigm::Int64; bign::Int64; bigt::Int64; bigk::Int64
    marketstartval::Float64; drift::Float64; marketvol::Float64
    impact_values::Ptr{Float64}; n_impact::Int64
    betamean::Float64; betastd::Float64; stockstartval::Float64
    vol_values::Ptr{Float64}; n_vol::Int64
    reviewprobability::Float64
    perf_lo::Int64; perf_hi::Int64
    cap_lo::Int64; cap_hi::Int64
    thresholdmean::Float64; thresholdstd::Float64
    psize_lo::Int64; psize_hi::Int64
end

# mirrors `struct fsb_config`
struct FsbConfig
    n_reps::Int64; rep_offset::Int64; reps_per_point::Int64
    device::Int32; keep_history::Int32; event_capacity::Int32; log_capacity::Int32
    trace_capacity::Int32; block_threads::Int32; blocks_per_sm::Int32
end

const SELECTORS = Dict("best" => 0, "goodenough" => 1, "probabilistic" => 2, "random" => 3)

lasterror() = unsafe_string(ccall((:fsb_last_error, LIB), Cstring, ()))
check(rc::Int32) = rc == 0 || error("libfsb_b200 error $rc: $(lasterror())")

mutable struct Engine
    handle::Ptr{Cvoid}
    keep::Vector{Any}          # arrays the POD points to
    function Engine(params; nreps=1, repoffset=0, device=0, keephistory=true)
        imp = collect(Float64, params.impactrange)
        vol = collect(Float64, params.stockvolrange)
        p = FsbParams(params.bigm, params.bign, params.bigt, params.bigk,
                      params.marketstartval, params.drift, params.marketvol,
                      pointer(imp), length(imp), params.betamean, params.betastd, params.stockstartval,
                      pointer(vol), length(vol), params.reviewprobability,
                      first(params.perfwindow), last(params.perfwindow),
                      params.invcaprange[1], params.invcaprange[2],
                      params.thresholdmean, params.thresholdstd,
                      first(params.portfsizerange), last(params.portfsizerange))
        cfg = FsbConfig(nreps, repoffset, 0, device, keephistory ? 1 : 0, 0, 0, 0, 0, 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve imp vol check(ccall((:fsb_create, LIB), Int32,
            (Ref{FsbParams}, Int32, Ref{FsbConfig}, Ref{Ptr{Cvoid}}), p, 1, cfg, h))
        e = new(h[], Any[imp, vol])
        finalizer(x -> (x.handle != C_NULL && ccall((:fsb_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle); x.handle = C_NULL), e)
        return e
    end
end

# Julia passes the struct fields directly: Matrix{Float64}/Vector{Float64}/Vector{Int64} are
# column-major and rooted for the duration of the ccall.
function upload!(e::Engine, market, stocks, investors, funds; rep=0)
    check(ccall((:fsb_upload_state, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        e.handle, rep, market.value, stocks.value, stocks.beta, stocks.vol, stocks.impact, stocks.volume,
        investors.assets, investors.horizon, investors.threshold, funds.holdings, funds.stakes, funds.value))
end

function download!(e::Engine, market, stocks, investors, funds; rep=0)
    check(ccall((:fsb_download_state, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        e.handle, rep, market.value, stocks.value, stocks.beta, stocks.vol, stocks.impact, stocks.volume,
        investors.assets, investors.horizon, investors.threshold, funds.holdings, funds.stakes, funds.value))
end

"Replacement for Func.initialise (src/functions.jl:417-430): Philox draws on the device."
function initialise(market, stocks, investors, funds, params; seed=0, rep=0, device=0)
    e = Engine(params; repoffset=rep, device=device)
    check(ccall((:fsb_init_device, LIB), Int32, (Ptr{Cvoid}, UInt64), e.handle, seed))
    download!(e, market, stocks, investors, funds)
    finalize(e)
end

"Replacement for Func.modelrun (src/functions.jl:446-491); returns the log as a NamedTuple of columns (:452-456)."
function modelrun(market, stocks, investors, funds, params, fundselector="bestperformer"; seed=0, rep=0, device=0,
                  stylefacts=nothing)
    haskey(SELECTORS, fundselector) || error("Please specify a valid fund selection method in reinvest!")  # Q9
    e = Engine(params; repoffset=rep, device=device)
    upload!(e, market, stocks, investors, funds)
    check(ccall((:fsb_set_seed, LIB), Int32, (Ptr{Cvoid}, UInt64), e.handle, seed))
    check(ccall((:fsb_run, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, Int32),
                e.handle, last(params.perfwindow) + 2, params.bigt, SELECTORS[fundselector], 0))
    download!(e, market, stocks, investors, funds)
    n = Ref{Int64}(0)
    check(ccall((:fsb_download_log, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
        e.handle, 0, n, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL))
    liquidated = zeros(Int64, n[]); replacedby = zeros(Int64, n[]); failtime = zeros(Int64, n[])
    ninvestors = zeros(Int64, n[]); nstocks = zeros(Int64, n[])
    initialsize = zeros(Float64, n[]); liquidity = zeros(Float64, n[])
    check(ccall((:fsb_download_log, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
        e.handle, 0, n, liquidated, replacedby, initialsize, liquidity, failtime, ninvestors, nstocks))
    # `stylefacts = Ref{Any}()`: receives calc_stylisedfacts of the finished run, computed from the histories
    # still on the device (src/runmodel.jl:47), before the engine is released
    stylefacts === nothing || (stylefacts[] = stylisedfacts(e, params))
    finalize(e)
    # wrap with DataFrame(...) where DataFrames is loaded: logisticregression (:846-856) and
    # plot_liquidationcounts (:832) read exactly these column names
    return (liquidated=liquidated, replacedby=replacedby, initialsize=initialsize, liquidity=liquidity,
            failtime=failtime, ninvestors=ninvestors, nstocks=nstocks)
end

# Func.calc_stylisedfacts (src/functions.jl:604-648; called at src/runmodel.jl:47) from the histories an
# engine holds on the device after a run (modelrun's `stylefacts` keyword calls it before releasing its
# engine): only the per-stock vectors cross the host link.  Returns the
# array-valued entries of the reference's result tuple; the z-tests / KS test (:629-641) run on them in
# Julia as before.
function stylisedfacts(e::Engine, params; rep=0, maxlag=10, returnspctile=5)
    M = params.bigm
    stockpacf = zeros(M, maxlag); stockvolacluster = zeros(M, maxlag)
    marketpacf = zeros(maxlag); marketvolacluster = zeros(maxlag)
    kurtoses = zeros(M); lossgainratios = zeros(M); corrs = zeros(M)
    check(ccall((:fsb_stylised_facts, LIB), Int32,
        (Ptr{Cvoid}, Int64, Int64, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        e.handle, rep, 1, maxlag, returnspctile, stockpacf, marketpacf, stockvolacluster, marketvolacluster,
        kurtoses, lossgainratios, corrs))
    nreturns = params.bigt - last(params.perfwindow) - 1
    return (stockpacf=(stockpacf, 1.96 / sqrt(M)), marketpacf=(marketpacf, 1.96 / sqrt(nreturns)),
            stockvolacluster=(stockvolacluster, 1.96 / sqrt(M)), marketvolacluster=(marketvolacluster, 1.96 / sqrt(nreturns)),
            stockkurtoses=kurtoses, lossgainratios=lossgainratios, corrs=corrs)
end

# The sweep of src/calibrate.jl:1-30 from ONE Julia task over several devices: every (point, replication) pair is a
# replication of an engine, engines hold contiguous shards of the flattened index (one per device), the days are enqueued
# asynchronously on every device, and the integer ensemble statistics are summed by one grouped allreduce
# (fsb_comm_init_all + fsb_stats_all).  Returns a (points x words) Matrix{UInt64}; layout: include/fsb.h "statistics".
function calibrate(points::AbstractVector, reps::Integer; devices=[0], seed=0, fundselector="probabilistic")
    haskey(SELECTORS, fundselector) || error("Please specify a valid fund selection method in reinvest!")
    keep = Any[]
    pods = map(points) do params
        imp = collect(Float64, params.impactrange); vol = collect(Float64, params.stockvolrange)
        push!(keep, imp, vol)
        FsbParams(params.bigm, params.bign, params.bigt, params.bigk, params.marketstartval, params.drift, params.marketvol,
                  pointer(imp), length(imp), params.betamean, params.betastd, params.stockstartval,
                  pointer(vol), length(vol), params.reviewprobability, first(params.perfwindow), last(params.perfwindow),
                  params.invcaprange[1], params.invcaprange[2], params.thresholdmean, params.thresholdstd,
                  first(params.portfsizerange), last(params.portfsizerange))
    end
    total = length(points) * reps
    ndev = length(devices)
    total >= ndev || error("$total (point, replication) pairs cannot be sharded over $ndev devices")
    handles = Ptr{Cvoid}[]
    GC.@preserve keep pods try
        for (g, dev) in enumerate(devices)
            lo, hi = div(total * (g - 1), ndev), div(total * g, ndev)
            cfg = FsbConfig(hi - lo, lo, reps, dev, 0, 0, 0, 0, 0, 0)
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:fsb_create, LIB), Int32, (Ptr{FsbParams}, Int32, Ref{FsbConfig}, Ref{Ptr{Cvoid}}), pods, length(pods), cfg, h))
            push!(handles, h[])
        end
        ndev > 1 && check(ccall((:fsb_comm_init_all, LIB), Int32, (Ptr{Ptr{Cvoid}}, Int32), handles, ndev))
        p1 = first(points)
        for h in handles
            check(ccall((:fsb_init_device, LIB), Int32, (Ptr{Cvoid}, UInt64), h, seed))
            check(ccall((:fsb_run_async, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, Int32),
                        h, last(p1.perfwindow) + 2, p1.bigt, SELECTORS[fundselector], 0))
        end
        foreach(h -> check(ccall((:fsb_sync, LIB), Int32, (Ptr{Cvoid},), h)), handles)
        words = ccall((:fsb_stats_words_per_point, LIB), Int64, (Ptr{Cvoid},), handles[1])
        out = zeros(UInt64, words, length(points))
        check(ccall((:fsb_stats_all, LIB), Int32, (Ptr{Ptr{Cvoid}}, Int32, Ptr{UInt64}), handles, ndev, out))
        return permutedims(out)
    finally
        foreach(h -> ccall((:fsb_destroy, LIB), Int32, (Ptr{Cvoid},), h), handles)
    end
end

end # module
