# B200.jl -- the reference-side shim a maintainer adds to Oceananigans (src/Architectures.jl and a new
# src/B200/ directory).  It cannot be executed in this image (no Julia toolchain); it is written against
# Oceananigans v0.76.8 and against include/ocean_b200.h.  The Python package ocean_b200 is the
# executable stand-in used by the tests.
#
# Everything is dispatch on the new architecture singleton; the reference's API is unchanged:
#     grid  = RectilinearGrid(B200(), size=(256,256,256), extent=(1,1,1), topology=(Periodic,Periodic,Periodic))
#     model = NonhydrostaticModel(; grid, advection=WENO5(), tracers=:b, buoyancy=BuoyancyTracer(),
#                                  timestepper=:RungeKutta3)
#     set!(model, u=..., v=...); run!(Simulation(model, Δt=..., stop_iteration=...))

module B200Arch

using Oceananigans
using Oceananigans.Architectures: AbstractArchitecture
using Oceananigans.Grids: RectilinearGrid, Periodic, Bounded, Flat, topology, halo_size
using Oceananigans.Fields: Field, location, boundary_conditions
using Oceananigans.BoundaryConditions: Flux, Value, Gradient, Open, Periodic as PeriodicBCClass
using Oceananigans.Models.NonhydrostaticModels: NonhydrostaticModel
import Oceananigans.Architectures: device, array_type, arch_array, architecture, device_event
import Oceananigans.TimeSteppers: time_step!, update_state!, calculate_tendencies!,
                                  calculate_pressure_correction!, pressure_correct_velocities!, store_tendencies!
import Oceananigans.BoundaryConditions: fill_halo_regions!
import Oceananigans.Solvers: solve!
import Oceananigans.Utils: launch!

const lib = "libocean_b200.so"

"The new architecture singleton (src/Architectures.jl:68-75)."
struct B200 <: AbstractArchitecture end

check(status::Int32) = status == 0 || begin
    n = ccall((:ob200_last_error, lib), Csize_t, (Ptr{UInt8}, Csize_t), C_NULL, 0)
    buf = Vector{UInt8}(undef, n + 1)
    ccall((:ob200_last_error, lib), Csize_t, (Ptr{UInt8}, Csize_t), buf, n + 1)
    error(unsafe_string(pointer(buf)))
end

# ---- Architectures.jl:81-143 ------------------------------------------------------------------
device(::B200) = (check(ccall((:ob200_init, lib), Int32, (Int32,), 0)); nothing)
launch!(::B200, args...; kw...) = error("B200(): KernelAbstractions kernels are not used on this architecture")
device_event(::B200) = nothing

"Device array owning a library allocation (array_type(::B200))."
mutable struct B200Array{T, N} <: AbstractArray{T, N}
    ptr  :: Ptr{Cvoid}
    dims :: NTuple{N, Int}
    function B200Array{T, N}(dims) where {T, N}
        p = Ref{Ptr{Cvoid}}()
        check(ccall((:ob200_malloc, lib), Int32, (Ref{Ptr{Cvoid}}, Csize_t), p, prod(dims) * sizeof(T)))
        a = new{T, N}(p[], dims)
        finalizer(x -> ccall((:ob200_free, lib), Int32, (Ptr{Cvoid},), x.ptr), a)
    end
end
Base.size(a::B200Array) = a.dims
array_type(::B200) = B200Array
architecture(::B200Array) = B200()
function arch_array(::B200, a::Array{T, N}) where {T, N}
    d = B200Array{T, N}(size(a))
    GC.@preserve a check(ccall((:ob200_upload, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), d.ptr, a, sizeof(a)))
    return d
end
function Base.Array(d::B200Array{T, N}) where {T, N}
    a = Array{T, N}(undef, d.dims)
    GC.@preserve a check(ccall((:ob200_download, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), a, d.ptr, sizeof(a)))
    return a
end

# ---- descriptors (mirror the POD structs of ocean_b200.h) ---------------------------------------
struct GridDesc
    ftype::Int32; N::NTuple{3,Int32}; H::NTuple{3,Int32}; topology::NTuple{3,Int32}
    L::NTuple{3,Float64}; regular::NTuple{3,Int32}; delta::NTuple{3,Float64}
    dC::NTuple{3,Ptr{Float64}}; dC_first::NTuple{3,Int32}; dC_len::NTuple{3,Int32}
    dF::NTuple{3,Ptr{Float64}}; dF_first::NTuple{3,Int32}; dF_len::NTuple{3,Int32}
end
topo_code(::Type{Periodic}) = Int32(0); topo_code(::Type{Bounded}) = Int32(1); topo_code(::Type{Flat}) = Int32(2)

"Marshal a RectilinearGrid (src/Grids/rectilinear_grid.jl:1-47) into an ob200_grid handle."
function grid_handle(grid::RectilinearGrid{FT}) where FT
    Δc = (grid.Δxᶜᵃᵃ, grid.Δyᵃᶜᵃ, grid.Δzᵃᵃᶜ); Δf = (grid.Δxᶠᵃᵃ, grid.Δyᵃᶠᵃ, grid.Δzᵃᵃᶠ)
    reg = map(x -> x isa Number, Δc)
    vec64(x) = x isa Number ? Float64[] : Float64.(parent(x))
    first_index(x) = x isa Number ? Int32(0) : Int32(first(axes(x, 1)))
    keep = (map(vec64, Δc), map(vec64, Δf))
    desc = GridDesc(FT == Float32 ? 0 : 1, Int32.(size(grid)), Int32.(halo_size(grid)), topo_code.(topology(grid)),
                    Float64.((grid.Lx, grid.Ly, grid.Lz)), Int32.(reg), map(x -> x isa Number ? Float64(x) : 0.0, Δc),
                    map(pointer, keep[1]), map(first_index, Δc), Int32.(map(length, keep[1])),
                    map(pointer, keep[2]), map(first_index, Δf), Int32.(map(length, keep[2])))
    h = Ref{Ptr{Cvoid}}()
    GC.@preserve keep check(ccall((:ob200_grid_create, lib), Int32, (Ref{GridDesc}, Ref{Ptr{Cvoid}}), desc, h))
    return h[]
end

# ---- model: nonhydrostatic_model.jl:102-203.  The constructor runs unchanged; on B200 its last step attaches a
# library model whose fields alias model.velocities / model.tracers / model.pressures / timestepper.Gⁿ, G⁻ ----------
const B200Model = NonhydrostaticModel{<:Any, <:Any, <:B200}
handle(model::B200Model) = model.auxiliary_fields.b200_handle     # set by the B200 method of the constructor

"Translate type parameters into the integer/struct configuration of ob200_model_desc; reject everything else."
function model_desc(model) end   # advection -> OB200_ADV_*, closure -> OB200_CLOSURE_*, FPlane, BuoyancyTracer,
                                 # constant boundary conditions -> ob200_bc; throws ArgumentError for function BCs,
                                 # forcings, background fields, LES closures, immersed grids, particles.

time_step!(model::B200Model, Δt; euler=false) =
    check(ccall((:ob200_model_time_step, lib), Int32, (Ptr{Cvoid}, Float64, Int32), handle(model), Δt, euler))
update_state!(model::B200Model) =
    check(ccall((:ob200_model_update_state, lib), Int32, (Ptr{Cvoid},), handle(model)))
calculate_tendencies!(model::B200Model) =
    check(ccall((:ob200_model_calculate_tendencies, lib), Int32, (Ptr{Cvoid},), handle(model)))
# calculate_pressure_correction! + pressure_correct_velocities! are one library call; the second is a no-op
calculate_pressure_correction!(model::B200Model, Δt) =
    check(ccall((:ob200_model_pressure_project, lib), Int32, (Ptr{Cvoid}, Float64), handle(model), Δt))
pressure_correct_velocities!(::B200Model, Δt) = nothing
store_tendencies!(::B200Model) = nothing            # pointer swap inside the library

# ---- fields and solvers ---------------------------------------------------------------------------
fill_halo_regions!(fields::NTuple{N, Field{<:Any,<:Any,<:Any,<:Any,<:RectilinearGrid{<:Any,<:Any,<:Any,<:Any,<:Any,<:Any,<:Any,<:Any,<:Any,<:Any,<:B200}}}, args...) where N =
    check(ccall((:ob200_fill_halo_regions, lib), Int32, (Ptr{Ptr{Cvoid}}, Int32), [f.data.handle for f in fields], N))

solve!(ϕ, solver::Oceananigans.Solvers.FFTBasedPoissonSolver{<:RectilinearGrid}, rhs) =
    GC.@preserve rhs check(ccall((:ob200_poisson_solve, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                                 solver.storage.handle, ϕ.data.handle, real.(Array(rhs))))

export B200
end # module
