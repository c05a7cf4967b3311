# record_tapes.jl -- equivalence-mode recorder (SURVEY.md 8c "tape", 8f4).
#
# Runs ONE replication of the UNMODIFIED reference (FundStabABM.jl: Func.initialise + the loop of
# Func.modelrun, src/functions.jl:417-430, 446-491, calling the reference's own functions in the
# reference's order) and writes, into one binary file,
#   * the state after Func.initialise,
#   * every random draw the time loop consumed, as the tapes libfsb_b200 / the oracle accept
#     (dense: z_mkt[T], z_stock[M x T], u_review[N x T]; events: anchors, portfolio sizes and picks,
#     fund choices),
#   * the state and the liquidation log after the last day.
# scripts/replay_reference_tapes.py uploads the first, feeds the second and compares with the third.
#
# How the draws are captured without touching the reference: before each call that consumes the
# global RNG its state is copied; after the call the SAME draws are replayed from the copy with the
# call shapes of the reference's call sites (src/functions.jl:15, 32, 149, 282, 123, 125, 392, 408,
# 413), and the copy must end in the state of the global RNG -- a mismatch aborts, so a tape is never
# written from a replay that consumed a different number of draws than the reference did.
# Fund choices are recorded as the index the reference actually chose (exact by construction) and,
# for the "probabilistic" selector, also as the uniform believed to drive Distributions.Categorical
# (one rand(); the replay tool can use either).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The file format is exercised
# from Python (fundstababm.jl_b200/tapefile.py, tests/) with the CPU oracle standing in as producer.
#
#   julia --project=<FundStabABM.jl checkout> record_tapes.jl out.fsbtape [seed=1] [selector=probabilistic]
#
# Needs bigk == bigm (the reference's own modelrun sizes its log with bigm, SURVEY.md Q4).

using Random
using DataFrames   # (a dependency of the reference's own project)
using FundStabABM
const Func = FundStabABM.Func
const Types = FundStabABM.Types
const Params = FundStabABM.Params

const SITE_ANCHOR, SITE_RPSIZE, SITE_RPPICK, SITE_CHOICE_U, SITE_CHOICE_IDX = 1, 2, 3, 4, 5
const SELECTORS = Dict("best" => 0, "goodenough" => 1, "probabilistic" => 2, "random" => 3)

globalrng() = isdefined(Random, :default_rng) ? Random.default_rng() : Random.GLOBAL_RNG
snapshot() = copy(globalrng())
function mustmatch(shadow, what)
    shadow == copy(globalrng()) || error("shadow replay of $what consumed different draws than the reference")
end

function writestate(io, market, stocks, investors, funds)
    write(io, market.value); write(io, stocks.value); write(io, stocks.beta); write(io, stocks.vol)
    write(io, stocks.impact); write(io, stocks.volume); write(io, investors.assets)
    write(io, Int64.(investors.horizon)); write(io, investors.threshold)
    write(io, funds.holdings); write(io, funds.stakes); write(io, funds.value)
end

function record(path; seed=1, selector="probabilistic", params=Params.default())
    bigk, bigm, bign, bigt = params.bigk, params.bigm, params.bign, params.bigt
    bigk == bigm || error("the reference's modelrun needs bigk == bigm (Q4)")
    W = params.perfwindow[end]
    Random.seed!(seed)
    market = Types.MarketIndex(zeros(bigt))
    stocks = Types.Equity(zeros(bigm, bigt), zeros(bigm), zeros(bigm), zeros(bigm), zeros(bigm, bigt))
    investors = Types.RetailInvestor(zeros(bign, bigk + 1), zeros(bign), zeros(bign))
    funds = Types.EquityFund(zeros(bigk, bigm), zeros(bigk, bign), zeros(bigk, bigt))
    Func.initialise(market, stocks, investors, funds, params)

    init = IOBuffer(); writestate(init, market, stocks, investors, funds)
    z_mkt = zeros(bigt); z_stock = zeros(bigm, bigt); u_review = zeros(bign, bigt)
    events = Tuple{Int64,Int64,Int64,Int64,Float64}[]   # (site, t, a, b, value), ids 0-based

    # the reference's log, exactly as src/functions.jl:448-456 builds it
    initialliquidity = [Func.calc_liquidity(funds, stocks, fundidx, 1) for fundidx in 1:bigm]
    liquidationlog = DataFrame(liquidated=zeros(Int64, bigm), replacedby=zeros(Int64, bigm),
        initialsize=copy(funds.value[:, 1]), liquidity=initialliquidity, failtime=zeros(Int64, bigm),
        ninvestors=dropdims(sum(funds.stakes .> 0, dims=2), dims=2),
        nstocks=dropdims(sum(funds.holdings .> 0, dims=2), dims=2))

    for t in W+2:bigt
        sh = snapshot(); Func.marketmove!(market.value, t, params)                    # :457
        z_mkt[t] = randn(sh); mustmatch(sh, "marketmove!")
        sh = snapshot(); Func.stockmove!(stocks.value, t, market.value, stocks.beta, stocks.vol)  # :458
        z_stock[:, t] .= randn(sh, bigm); mustmatch(sh, "stockmove!")
        Func.fundrevalue!(funds, stocks.value, 1:bigk)                                 # :459
        sh = snapshot(); reviewers = Func.drawreviewers(params)                        # :461
        u_review[:, t] .= rand(sh, bign); mustmatch(sh, "drawreviewers")
        divestments = Func.perfreview(t, reviewers, investors, funds.value)            # :462
        if sum(divestments[1, :]) > 0
            sellorders, _, _ = Func.liquidate!(funds.holdings, funds.stakes, stocks.value[:, t], divestments)
            Func.trackvolume!(stocks.volume, sellorders, t)
            _, cashout = Func.executeorder!(stocks, t, sellorders)
            Func.disburse!(investors.assets, divestments, cashout)
            if any(sum(funds.stakes, dims=2) .== 0)                                    # :471
                spawn = findall(vec(sum(funds.holdings, dims=2) .== 0))                # what respawn! will see (:279-281)
                potentialinvs = findall(vec(investors.assets[:, end] .> 0))
                sh = snapshot()
                buyorders, _, _, liquidationlog = Func.respawn!(funds, investors, t, stocks, params, liquidationlog)
                anchors = Random.shuffle(sh, potentialinvs)[1:length(spawn)]           # :282
                for idx in 1:length(spawn)
                    push!(events, (SITE_ANCHOR, t, idx - 1, 0, Float64(anchors[idx] - 1)))
                    nofm = rand(sh, params.portfsizerange, 1)                          # :123 (one fund row)
                    push!(events, (SITE_RPSIZE, t, idx - 1, 0, Float64(nofm[1])))
                    selection = rand(sh, 1:bigm, nofm[1])                              # :125
                    for (j, s) in enumerate(selection)
                        push!(events, (SITE_RPPICK, t, idx - 1, j - 1, Float64(s - 1)))
                    end
                end
                mustmatch(sh, "respawn!")
                Func.trackvolume!(stocks.volume, sellorders, t)
                _, sharesout = Func.executeorder!(stocks, t, buyorders)
                Func.disburse!(funds, sharesout, stocks.value)
            end
            Func.fundrevalue!(funds, stocks.value, 1:bigk)                             # :479
            if any(investors.assets[:, end] .> 0)
                reinvestors = findall(vec(investors.assets[:, end] .> 0))              # :339
                sh = snapshot()
                _, _, buyorders = Func.reinvest!(investors, funds, stocks.value, t, selector)
                for (i, reinv) in enumerate(reinvestors)                               # the choice the reference made
                    push!(events, (SITE_CHOICE_IDX, t, reinv - 1, 0, Float64(buyorders.funds[i] - 1)))
                end
                if selector == "probabilistic"                                         # :407-408, believed: one rand() each
                    for reinv in reinvestors
                        push!(events, (SITE_CHOICE_U, t, reinv - 1, 0, rand(sh)))
                    end
                    sh == copy(globalrng()) || @warn "Categorical sampling does not consume one rand() per choice: CHOICE_U records are not usable" t
                end
                Func.trackvolume!(stocks.volume, sellorders, t)
                _, sharesout = Func.executeorder!(stocks, t, buyorders)
                Func.disburse!(funds, sharesout, stocks.value)
            end
        end
    end

    nlog = size(liquidationlog, 1)
    open(path, "w") do io
        write(io, "FSBTAPE1")
        write(io, Int64[bigk, bigm, bign, bigt, W, SELECTORS[selector], length(events), nlog, 1, seed,
                        reinterpret(Int64, Float64(params.reviewprobability)), 0])
        write(io, take!(init))
        write(io, z_mkt); write(io, z_stock); write(io, u_review)
        for (site, t, a, b, v) in events
            write(io, Int64[site, t, a, b]); write(io, Float64(v))
        end
        writestate(io, market, stocks, investors, funds)
        write(io, Int64.(liquidationlog.liquidated)); write(io, Int64.(liquidationlog.replacedby))
        write(io, Float64.(liquidationlog.initialsize)); write(io, Float64.(liquidationlog.liquidity))
        write(io, Int64.(liquidationlog.failtime)); write(io, Int64.(liquidationlog.ninvestors))
        write(io, Int64.(liquidationlog.nstocks))
    end
    println("wrote ", path, ": ", length(events), " event draws, ", nlog, " log rows")
end

if abspath(PROGRAM_FILE) == @__FILE__
    length(ARGS) >= 1 || error("usage: record_tapes.jl out.fsbtape [seed] [selector]")
    record(ARGS[1]; seed=length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 1,
           selector=length(ARGS) >= 3 ? ARGS[3] : "probabilistic")
end
