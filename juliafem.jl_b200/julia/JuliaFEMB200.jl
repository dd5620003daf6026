# JuliaFEMB200.jl -- thin `ccall` shim that plugs libjfem_b200.so (include/jfem_b200.h) into JuliaFEM.jl.
#
# It replaces the ext/JuliaFEMCUDAExt.jl route: the two generic functions the reference dispatches on,
#     initialize_backend(::GPU, physics, time)          (src/backend/abstract.jl:228, ext:867-869)
#     solve_backend!(data, physics; tol, max_iter, ...)  (src/backend/abstract.jl:241, ext:886-916)
# get methods for a new data type `ElasticityDataB200`, and the classic path gets an `assemble_elements!` method
# (the documented override point, src/assembly/assembly.jl:19-29) that fills `assembly.K` / `assembly.f` from the device.
#
# NOTE: no Julia toolchain exists in the build image, so this file has been reviewed but never executed.  Every call
# below is a 1:1 binding of a C-ABI function that IS exercised by the Python host (juliafem.jl_b200/_lib.py) and the
# GPU test-suite; argument order and types follow include/jfem_b200.h.
module JuliaFEMB200

using JuliaFEM
import JuliaFEM: initialize_backend, solve_backend!, assemble_elements!, GPU, Physics, Problem, Elasticity, Assembly, Element,
                 AbstractElasticityData, add!

const libjfem = get(ENV, "JFEM_B200_LIB", joinpath(@__DIR__, "..", "libjfem_b200.so"))

const JFEM_TET4, JFEM_HEX8, JFEM_TET10 = Cint(4), Cint(8), Cint(10)
const JFEM_MAT_LINEAR_ELASTIC, JFEM_MAT_NEO_HOOKEAN, JFEM_MAT_PERFECT_PLASTICITY = Cint(0), Cint(1), Cint(2)
const JFEM_PROJECT, JFEM_TANGENT, JFEM_USE_CSR = Cint(1), Cint(2), Cint(4)

last_error() = unsafe_string(ccall((:jfem_last_error, libjfem), Cstring, ()))

"Turn a non-zero status into a Julia exception, as the reference does with `error(...)` / `DomainError`."
function check(rc::Cint)
    rc == 0 && return nothing
    rc == 5 && throw(DomainError(NaN, last_error()))      # J <= 0, src/materials/neo_hookean.jl:137
    error("libjfem_b200 error $rc: $(last_error())")
end

mutable struct ElasticityDataB200 <: AbstractElasticityData
    handle::Ptr{Cvoid}
    node_ids::Vector{Int}          # sorted unique original ids (ext:100-108)
    n_dofs::Int
    f_ext::Vector{Float64}
    prescribed::Vector{Float64}
    function ElasticityDataB200(handle, node_ids, n_dofs)
        d = new(handle, node_ids, n_dofs, zeros(n_dofs), zeros(n_dofs))
        finalizer(x -> (x.handle != C_NULL && ccall((:jfem_destroy, libjfem), Cint, (Ptr{Cvoid},), x.handle); x.handle = C_NULL), d)
        return d
    end
end

nnpe_code(::Type{JuliaFEM.Tet4}) = JFEM_TET4
nnpe_code(::Type{JuliaFEM.Hex8}) = JFEM_HEX8
nnpe_code(::Type{JuliaFEM.Tet10}) = JFEM_TET10

"Flatten `physics.body_elements` into the arrays jfem_create takes (same walk as initialize_gpu_data!, ext:86-217)."
function create_handle(elements::Vector, device::Integer)
    T = typeof(first(elements)).parameters[1]               # topology type parameter of Element{...}
    nn = Int(nnpe_code(T))
    conn_orig = [Int.(collect(el.connectivity)) for el in elements]
    node_ids = sort(unique(vcat(conn_orig...)))
    remap = Dict(id => i for (i, id) in enumerate(node_ids))
    coords = zeros(Float64, 3, length(node_ids))
    conn = zeros(Int32, nn, length(elements))
    for (e, el) in enumerate(elements)
        X = el.fields.geometry                               # 3 x nn
        for k in 1:nn
            j = remap[conn_orig[e][k]]
            conn[k, e] = j
            coords[:, j] .= X[:, k]
        end
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve coords conn begin
        check(ccall((:jfem_create, libjfem), Cint,
                    (Ref{Ptr{Cvoid}}, Cint, Cint, Int64, Int64, Ptr{Float64}, Ptr{Int32}, Cint),
                    h, device, nnpe_code(T), length(node_ids), length(elements), coords, conn, 1))
    end
    return h[], node_ids, remap
end

"Tuning knobs of the library (include/jfem_b200.h: \"patch_elems\", \"deterministic\", \"affine_fast_path\", \"warp_specialised\", ...)."
set_option!(data::ElasticityDataB200, key::AbstractString, value::Real) =
    check(ccall((:jfem_set_option, libjfem), Cint, (Ptr{Cvoid}, Cstring, Cdouble), data.handle, key, Float64(value)))

function initialize_backend(backend::GPU, physics::Physics, time::Float64)
    handle, node_ids, remap = create_handle(physics.body_elements, 0)
    data = ElasticityDataB200(handle, node_ids, 3 * length(node_ids))
    el1 = first(physics.body_elements)
    params = Float64[el1.fields.youngs_modulus, el1.fields.poissons_ratio]
    kind = physics.properties.finite_strain ? JFEM_MAT_NEO_HOOKEAN : JFEM_MAT_LINEAR_ELASTIC
    check(ccall((:jfem_set_material, libjfem), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Cint), handle, kind, params, 2, 0))
    dofs = Int64[]; vals = Float64[]
    bc = physics.bc_dirichlet
    for (nodes, comps, val) in zip(bc.node_ids, bc.components, bc.values)
        for n in nodes, c in comps
            push!(dofs, 3 * (remap[n] - 1) + c); push!(vals, val)
        end
    end
    check(ccall((:jfem_set_dirichlet, libjfem), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Int64), handle, dofs, vals, length(dofs)))
    data.prescribed[dofs] .= vals
    # Tri3 lumped traction, area/3 * t per node (apply_surface_traction_kernel!, ext:368-416) -- host side, O(surface)
    for (surf, t) in zip(physics.bc_neumann.surface_elements, physics.bc_neumann.traction)
        X = surf.fields.geometry
        a = 0.5 * sqrt(sum(abs2, JuliaFEM.cross(X[:, 2] - X[:, 1], X[:, 3] - X[:, 1])))
        for n in surf.connectivity, c in 1:3
            data.f_ext[3 * (remap[Int(n)] - 1) + c] += a / 3 * t[c]
        end
    end
    return data
end

function matvec(data::ElasticityDataB200, x::Vector{Float64}; flags::Cint=Cint(0))
    y = similar(x)
    check(ccall((:jfem_matvec, libjfem), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Cint), data.handle, x, y, flags, 0))
    return y
end

function solve_backend!(data::ElasticityDataB200, physics::Physics; tol=1e-6, max_iter=1000, newton_tol=1e-6, max_newton=20,
                        max_cg_per_newton=50)
    if physics.properties.finite_strain
        u = copy(data.prescribed)
        nit = Ref{Cint}(0); cgit = Ref{Cint}(0); res = Ref{Float64}(0.0)
        hist = zeros(Float64, 3, max_newton + 1)
        check(ccall((:jfem_newton_krylov, libjfem), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Cint, Cint, Float64, Float64, Cint, Ref{Cint}, Ref{Cint},
                     Ref{Float64}, Ptr{Float64}, Cint, Cint),
                    data.handle, data.f_ext, u, newton_tol, max_newton, max_cg_per_newton, 0.5, 0.9, 0, nit, cgit, res, hist,
                    max_newton + 1, 0))
        history = [(Int(hist[1, k]), hist[2, k], hist[3, k]) for k in 1:nit[]]
        return (u, Int(nit[]), Int(cgit[]), res[], history)
    end
    b = any(!iszero, data.prescribed) ? data.f_ext - matvec(data, data.prescribed) : copy(data.f_ext)
    x = zeros(data.n_dofs)
    it = Ref{Cint}(0); res = Ref{Float64}(0.0)
    check(ccall((:jfem_cg, libjfem), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Cint, Cint, Cint, Ref{Cint}, Ref{Float64}, Cint),
                data.handle, b, x, tol, 0, max_iter, 0, it, res, 0))
    it[] == max_iter && @warn "CG did not converge in $max_iter iterations"        # ext:630
    return (x + data.prescribed, 1, Int(it[]), res[], [(Int(it[]), sqrt(sum(abs2, b)), 0.0)])
end

"""
Classic path: `assemble!(problem::Problem{Elasticity}, time)` ends in `assemble_elements!` (src/assembly/assembly.jl:25);
this method computes K on the GPU and appends it to `assembly.K` as COO triplets so that `solve!(analysis)` works unchanged.
"""
function assemble_elements!(problem::Problem{Elasticity}, assembly::Assembly, elements::Vector{Element{T}}, time) where T
    handle, node_ids, remap = create_handle(elements, 0)
    try
        el1 = first(elements)
        params = Float64[el1("youngs modulus", time), el1("poissons ratio", time)]
        check(ccall((:jfem_set_material, libjfem), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Cint), handle, JFEM_MAT_LINEAR_ELASTIC, params, 2, 0))
        n = Ref{Int64}(0); nnz = Ref{Int64}(0)
        check(ccall((:jfem_csr_size, libjfem), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), handle, n, nnz))
        rowptr = zeros(Int64, n[] + 1); colind = zeros(Int32, nnz[]); vals = zeros(Float64, nnz[])
        check(ccall((:jfem_csr_pattern, libjfem), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}), handle, rowptr, colind))
        check(ccall((:jfem_assemble_csr, libjfem), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint),
                    handle, C_NULL, vals, C_NULL, 0, 0))
        # local dense numbering -> the problem's dofs 3*(node-1)+c (src/assembly/problems.jl:476)
        gd(ld) = 3 * (node_ids[(ld - 1) ÷ 3 + 1] - 1) + (ld - 1) % 3 + 1
        for r in 1:n[], p in rowptr[r]:(rowptr[r + 1] - 1)
            add!(assembly.K, gd(r), gd(Int(colind[p])), vals[p])
        end
    finally
        ccall((:jfem_destroy, libjfem), Cint, (Ptr{Cvoid},), handle)
    end
    return nothing
end

end # module
