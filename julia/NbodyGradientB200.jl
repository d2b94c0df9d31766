# NbodyGradientB200.jl — Julia host side of libnbgrad_b200.so (include/nbgrad.h).
#
# Keeps the reference API (ElementsIC / CartesianIC -> State -> Integrator(h, t0, tmax) -> TransitTiming /
# TransitParameters, grad = true/false) and adds BATCH methods on the reference's own types that reach CUDA through one
# `ccall` per batch.  The integrator, Jacobian and transit subsystems run in hand-written sm_100a kernels; nothing on
# the path is computed in Julia and there is no CPU fallback (a missing library or device throws).
#
#   b200(intr)(ss::Vector{State}, tts::Vector{TransitTiming|TransitParameters}; grad)   <-> Transits.jl:140-180
#   b200(intr)(ss::Vector{State}, time::Float64; grad)                                  <-> Integrator.jl:159-197
#   b200(intr)(ss::Vector{State}, N::Int; grad)                                         <-> Integrator.jl:211-234
#   b200(intr)(ss::Vector{State}; grad)                                                 <-> Integrator.jl:247
#   single-system forms b200(intr)(s, tt) etc. are batches of one.
#
# NOTE: Julia is not available in the build environment of this repository, so this file is exercised only where a
# Julia toolchain exists; the same C ABI is exercised by the Python mirror (nbodygradient.jl_b200/nbgrad) in tests/.
module NbodyGradientB200

using NbodyGradient
using LinearAlgebra: I
import NbodyGradient: Integrator, State, TransitTiming, TransitParameters, TransitOutput, check_step

export b200, B200Integrator, nbg_device_count, chi2_fused

const LIB = get(ENV, "NBGRAD_B200_LIB", "libnbgrad_b200")

const NBG_ERRORS = Dict(-1 => "NBG_ERR_ARG", -2 => "NBG_ERR_NO_DEVICE", -3 => "NBG_ERR_CUDA", -4 => "NBG_ERR_UNSUPPORTED", -5 => "NBG_ERR_NOMEM")

struct NbgError <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::NbgError) = print(io, get(NBG_ERRORS, Int(e.code), string(e.code)), ": ", e.msg)

@inline function chk(rc::Int32)
    rc == 0 && return
    throw(NbgError(rc, unsafe_string(ccall((:nbg_last_error, LIB), Cstring, ()))))
end

nbg_device_count() = Int(ccall((:nbg_device_count, LIB), Int32, ()))

# ---- plan cache: (nbody, nsys, devices) -> nbg_plan*; bounded (a plan keeps its device buffers), oldest entry evicted ------------
const PLANS = Dict{Tuple{Int,Int,Vector{Int32}},Ptr{Cvoid}}()
const PLAN_ORDER = Tuple{Int,Int,Vector{Int32}}[]
const MAX_PLANS = 4
function plan(n::Int, nsys::Int, devices::Vector{Int32})
    key = (n, nsys, devices)
    haskey(PLANS, key) && return PLANS[key]
    while length(PLAN_ORDER) >= MAX_PLANS
        old = popfirst!(PLAN_ORDER)
        ccall((:nbg_plan_destroy, LIB), Int32, (Ptr{Cvoid},), pop!(PLANS, old))
    end
    p = Ref{Ptr{Cvoid}}(C_NULL)
    # one device: nbg_plan_create; several: one contiguous slice of the batch, one child plan and one host thread per device
    # inside the library (nbg_plan_create_multi) -- every call below then runs on all devices concurrently
    chk(ccall((:nbg_plan_create_multi, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32, Int64, Ptr{Int32}, Int32, Int64), p, n, nsys, devices, length(devices), 0))
    PLANS[key] = p[]
    push!(PLAN_ORDER, key)
    return p[]
end
function release_plans()
    for p in values(PLANS)
        ccall((:nbg_plan_destroy, LIB), Int32, (Ptr{Cvoid},), p)
    end
    empty!(PLANS); empty!(PLAN_ORDER)
end
atexit(release_plans)

"""
    b200(intr::Integrator; device=0, devices=[device])

Wrap a reference `Integrator` so that calling it runs on the B200(s).  `devices = 0:7` shards a batch over the eight GPUs of a box
(contiguous slices, no collective; SURVEY 8(e)).  `h`, `tmax` are read ONCE, by value (the reference mutates `intr.h` inside
`(intr)(s,N)`, Integrator.jl:218,232, so an Integrator must not be shared across threads).
"""
struct B200Integrator
    h::Float64
    t0::Float64
    tmax::Float64
    device::Vector{Int32}
end
b200(intr::Integrator; device::Int=0, devices=[device]) = B200Integrator(intr.h, intr.t0, intr.tmax, collect(Int32, devices))

# ---- packing: Julia column-major arrays with the system index slowest are exactly the ABI layout ----------------------
function pack(ss::Vector{State{Float64}})
    B, n = length(ss), ss[1].n
    all(s -> s.n == n, ss) || throw(ArgumentError("all systems of a batch must have the same number of bodies"))
    all(s -> s.t[1] == ss[1].t[1], ss) || throw(ArgumentError("all systems of a batch must share s.t"))
    all(s -> s.pair == ss[1].pair, ss) || throw(ArgumentError("all systems of a batch must share s.pair"))
    x = Array{Float64}(undef, 3, n, B); v = similar(x); xe = similar(x); ve = similar(x)
    m = Array{Float64}(undef, n, B)
    for (b, s) in enumerate(ss)
        x[:, :, b] .= s.x; v[:, :, b] .= s.v; xe[:, :, b] .= s.xerror; ve[:, :, b] .= s.verror; m[:, b] .= s.m
    end
    return x, v, m, xe, ve
end

function upload(p, ss::Vector{State{Float64}}, grad::Bool)
    x, v, m, xe, ve = pack(ss)
    # s.pair: Matrix{Bool} is one byte per entry in column-major order, exactly what nbg_set_pair reads
    chk(ccall((:nbg_set_pair, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), p, any(ss[1].pair) ? reinterpret(UInt8, ss[1].pair) : C_NULL))
    B, M = length(ss), 7 * ss[1].n
    if grad
        js = Array{Float64}(undef, M, M, B); je = similar(js); dq = Array{Float64}(undef, M, B)
        for (b, s) in enumerate(ss)
            js[:, :, b] .= s.jac_step; je[:, :, b] .= s.jac_error; dq[:, b] .= s.dqdt
        end
        chk(ccall((:nbg_set_state, LIB), Int32,
                  (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                  p, x, v, m, ss[1].t[1], xe, ve, js, je, dq))
    else
        chk(ccall((:nbg_set_state, LIB), Int32,
                  (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                  p, x, v, m, ss[1].t[1], xe, ve, C_NULL, C_NULL, C_NULL))
    end
end

# host buffers for the final state of a batch (ABI layout) and their scatter back into the reference's State objects
function state_buffers(B::Int, n::Int, grad::Bool)
    M = 7n
    x = Array{Float64}(undef, 3, n, B); v = similar(x); xe = similar(x); ve = similar(x)
    t = Vector{Float64}(undef, B); status = Vector{UInt32}(undef, B)
    js = grad ? Array{Float64}(undef, M, M, B) : nothing
    je = grad ? Array{Float64}(undef, M, M, B) : nothing
    dq = grad ? Array{Float64}(undef, M, B) : nothing
    return (x=x, v=v, xe=xe, ve=ve, js=js, je=je, dq=dq, t=t, status=status)
end
function scatter!(ss::Vector{State{Float64}}, o, grad::Bool)
    for (b, s) in enumerate(ss)
        s.x .= @view o.x[:, :, b]; s.v .= @view o.v[:, :, b]; s.xerror .= @view o.xe[:, :, b]; s.verror .= @view o.ve[:, :, b]
        s.t[1] = o.t[b]
        if grad
            s.jac_step .= @view o.js[:, :, b]; s.jac_error .= @view o.je[:, :, b]; s.dqdt .= @view o.dq[:, b]
        end
    end
    return o.status
end

function download!(p, ss::Vector{State{Float64}}, grad::Bool)
    o = state_buffers(length(ss), ss[1].n, grad)
    chk(ccall((:nbg_get_state, LIB), Int32,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt32}),
              p, o.x, o.v, o.xe, o.ve, grad ? o.js : C_NULL, grad ? o.je : C_NULL, grad ? o.dq : C_NULL, o.t, o.status))
    return scatter!(ss, o, grad)
end

# A State that no integrator call has touched: jac_step = I, error terms and dq/dh zero (what State(ic) constructs, Integrator.jl:82-103)
isfresh(s::State{Float64}) = all(iszero, s.xerror) && all(iszero, s.verror) && all(iszero, s.jac_error) && all(iszero, s.dqdt) && s.jac_step == I

# ---- (intr)(s, time; grad)  Integrator.jl:159-197 ----------------------------------------------------------------------
function (bi::B200Integrator)(ss::Vector{State{Float64}}, time::Float64; grad::Bool=true)
    t0 = ss[1].t[1]
    nsteps = abs(round(Int64, (time - t0) / bi.h))
    h = bi.h * check_step(t0, time)
    tmax = t0 + (h * nsteps)
    h_last = tmax != time ? time - tmax : 0.0
    p = plan(ss[1].n, length(ss), bi.device)
    upload(p, ss, grad)
    chk(ccall((:nbg_integrate_resident, LIB), Int32, (Ptr{Cvoid}, Float64, Int64, Float64, Int32, Int32, Float64), p, h, nsteps, h_last, grad, 1, time))
    return download!(p, ss, grad)
end

# ---- (intr)(s, N; grad)  Integrator.jl:211-234 --------------------------------------------------------------------------
function (bi::B200Integrator)(ss::Vector{State{Float64}}, N::Int64; grad::Bool=true)
    h = N < 0 ? -bi.h : bi.h
    p = plan(ss[1].n, length(ss), bi.device)
    upload(p, ss, grad)
    chk(ccall((:nbg_integrate_resident, LIB), Int32, (Ptr{Cvoid}, Float64, Int64, Float64, Int32, Int32, Float64), p, h, abs(N), 0.0, grad, 0, 0.0))
    return download!(p, ss, grad)
end

# ---- (intr)(s; grad)  Integrator.jl:247 -------------------------------------------------------------------------------
(bi::B200Integrator)(ss::Vector{State{Float64}}; grad::Bool=true) = bi(ss, ss[1].t[1] + bi.tmax; grad=grad)

# ---- (intr)(s, tt; grad)  Transits.jl:140-180 -------------------------------------------------------------------------
ncomp(::TransitTiming) = 1
ncomp(::TransitParameters) = 3

function (bi::B200Integrator)(ss::Vector{State{Float64}}, tts::Vector{<:TransitOutput{Float64}}; grad::Bool=true)
    B, n = length(ss), ss[1].n
    length(tts) == B || throw(ArgumentError("one transit output per system"))
    M, ntt, ti, C = 7n, tts[1].ntt, tts[1].ti, ncomp(tts[1])
    all(t -> t.ntt == ntt && t.ti == ti, tts) || throw(ArgumentError("all transit outputs of a batch must share ntt and ti"))
    p = plan(n, B, bi.device)
    ntt_body = fill(Int32(ntt), n)
    jinit = C_NULL
    ji = nothing
    if grad
        ji = Array{Float64}(undef, M, M, B)
        for (b, s) in enumerate(ss); ji[:, :, b] .= s.jac_init; end
    end
    # ABI layout (C order): tt[sys][i][k][c], dtdq0[sys][i][k][p][q][c]  ==  Julia arrays (c, k, i, b) and (c, q, p, k, i, b)
    traw = zeros(Float64, C, ntt, n, B)
    count = zeros(Int64, n, B)
    draw = grad ? zeros(Float64, C, 7, n, ntt, n, B) : nothing
    eraw = grad ? zeros(Float64, C, 7, n, ntt, n, B) : nothing
    out = nothing
    if all(isfresh, ss)
        # fresh State(ic) objects: the one-shot entry point takes the host arrays for inputs AND outputs, so the library streams every
        # chunk's transit rows into them while the next chunk computes and never holds the full arrays on the device (INTEGRATION.md 6)
        x, v, m, _, _ = pack(ss)
        out = state_buffers(B, n, grad)
        pairarg = any(ss[1].pair) ? reinterpret(UInt8, ss[1].pair) : C_NULL
        chk(ccall((:nbg_transit_timing, LIB), Int32,
                  (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Float64, Float64, Float64, Int32, Ptr{Int32}, Int32, Int32,
                   Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                   Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt32}),
                  p, x, v, m, pairarg, ss[1].t[1], bi.h, bi.tmax, ti - 1, ntt_body, C == 3 ? 1 : 0, grad,
                  grad ? ji : jinit, traw, count, grad ? draw : C_NULL, grad ? eraw : C_NULL, out.x, out.v, out.xe, out.ve,
                  grad ? out.js : C_NULL, grad ? out.je : C_NULL, grad ? out.dq : C_NULL, out.t, out.status))
    else
        upload(p, ss, grad)
        chk(ccall((:nbg_transit_timing_resident, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, Int32, Ptr{Int32}, Int32, Int32, Ptr{Float64}),
                  p, bi.h, bi.tmax, ti - 1, ntt_body, C == 3 ? 1 : 0, grad, grad ? ji : jinit))
        chk(ccall((:nbg_transit_fetch, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                  p, traw, count, grad ? draw : C_NULL, grad ? eraw : C_NULL))
    end
    for (b, tt) in enumerate(tts)
        tt.count .= @view count[:, b]
        if C == 1
            tt.tt .= permutedims(@view(traw[1, :, :, b]), (2, 1))                                  # [i,k]
            if grad
                tt.dtdq0 .= permutedims(@view(draw[1, :, :, :, :, b]), (4, 3, 1, 2))               # (q,p,k,i) -> [i,k,q,p]
                tt.dtdelements .= permutedims(@view(eraw[1, :, :, :, :, b]), (4, 3, 1, 2))
            end
        else
            tt.ttbv .= permutedims(@view(traw[:, :, :, b]), (1, 3, 2))                             # (c,k,i) -> [c,i,k]
            if grad
                tt.dtbvdq0 .= permutedims(@view(draw[:, :, :, :, :, b]), (1, 5, 4, 2, 3))          # (c,q,p,k,i) -> [c,i,k,q,p]
                tt.dtbvdelements .= permutedims(@view(eraw[:, :, :, :, :, b]), (1, 5, 4, 2, 3))
            end
        end
    end
    return out === nothing ? download!(p, ss, grad) : scatter!(ss, out, grad)
end

"""
    chi2_fused(bi, ss, tts, t_obs, sigma; wrt_elements=false)

The transit-timing run with the likelihood fused into the Jacobian kernel (nbg_transit_chi2_fused): returns `chi2[b]` and
`grad[q,p,b]` = d chi2 / d q0 (or d / d elements when the states were built on the device; not reachable from a host-built State).
`t_obs`, `sigma`: `ntt x n` (Julia `tt.tt'` order: slot k fastest within body i) shared by the batch.  No dtdq0 array exists anywhere.
"""
function chi2_fused(bi::B200Integrator, ss::Vector{State{Float64}}, tts::Vector{<:TransitTiming{Float64}}, t_obs::Matrix{Float64}, sigma::Matrix{Float64})
    B, n = length(ss), ss[1].n
    ntt, ti = tts[1].ntt, tts[1].ti
    p = plan(n, B, bi.device)
    upload(p, ss, true)
    chi2 = zeros(Float64, B); g = zeros(Float64, 7, n, B); count = zeros(Int64, n, B)
    chk(ccall((:nbg_transit_chi2_fused, LIB), Int32,
              (Ptr{Cvoid}, Float64, Float64, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
              p, bi.h, bi.tmax, ti - 1, fill(Int32(ntt), n), t_obs, sigma, 0, 0, 1, chi2, g, count, C_NULL))
    for (b, tt) in enumerate(tts); tt.count .= @view count[:, b]; end
    download!(p, ss, true)
    return chi2, g
end

# ---- (intr)(s, o::CartesianOutput)  src/outputs/Outputs.jl:26-49 --------------------------------------------------------
# o.states[i] = deepcopy(s) before step i: positions, velocities and jac_step of every saved state come back from ONE call
# (nbg_integrate_sampled_jac); the other fields of the copies (m, pair, n, ...) are those of the input state, t = t0 + h (i-1).
function (bi::B200Integrator)(ss::Vector{State{Float64}}, os::Vector{CartesianOutput{Float64}})
    B, n = length(ss), ss[1].n
    length(os) == B || throw(ArgumentError("one CartesianOutput per system"))
    nstep = os[1].nstep
    all(o -> o.nstep == nstep, os) || throw(ArgumentError("the batch shares nstep"))
    t0 = ss[1].t[1]
    h = bi.h * NbodyGradient.check_step(t0, bi.tmax)
    p = plan(n, B, bi.device)
    upload(p, ss, true)
    xs = zeros(Float64, 3, n, B, nstep); vs = zeros(Float64, 3, n, B, nstep); js = zeros(Float64, 7n, 7n, B, nstep)
    chk(ccall((:nbg_integrate_sampled_jac, LIB), Int32, (Ptr{Cvoid}, Float64, Int64, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              p, h, nstep, 1, 1, xs, vs, js))
    for (b, (s, o)) in enumerate(zip(ss, os)), i in 1:nstep
        c = deepcopy(s)
        c.x .= @view xs[:, :, b, i]; c.v .= @view vs[:, :, b, i]; c.jac_step .= @view js[:, :, b, i]
        c.t[1] = t0 + h * (i - 1)
        o.states[i] = c
    end
    download!(p, ss, true)
    for s in ss; s.t[1] = t0 + h * nstep; end
    return
end
(bi::B200Integrator)(s::State{Float64}, o::CartesianOutput{Float64}) = bi([s], [o])

# single-system forms: a batch of one
(bi::B200Integrator)(s::State{Float64}, tt::TransitOutput{Float64}; grad::Bool=true) = (bi([s], [tt]; grad=grad); nothing)
(bi::B200Integrator)(s::State{Float64}, time::Float64; grad::Bool=true) = (bi([s], time; grad=grad); nothing)
(bi::B200Integrator)(s::State{Float64}, N::Int64; grad::Bool=true) = (bi([s], N; grad=grad); nothing)
(bi::B200Integrator)(s::State{Float64}; grad::Bool=true) = (bi([s]; grad=grad); nothing)

end # module
