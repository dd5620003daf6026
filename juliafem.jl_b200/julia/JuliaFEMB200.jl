# JuliaFEMB200.jl -- thin `ccall` shim that plugs libjfem_b200.so (include/jfem_b200.h) into JuliaFEM.jl.
#
# It replaces the ext/JuliaFEMCUDAExt.jl route: the two generic functions the reference dispatches on,
#     initialize_backend(::GPU, physics, time)          (src/backend/abstract.jl:228, ext:867-869)
#     solve_backend!(data, physics; tol, max_iter, ...)  (src/backend/abstract.jl:241, ext:886-916)
# get methods for a new data type `ElasticityDataB200`, and the classic path gets an `assemble_elements!` method
# (the documented override point, src/assembly/assembly.jl:19-29) that fills `assembly.K` / `assembly.f` from the device.
#
# NOTE: no Julia toolchain exists in the build image, so this file has been reviewed but never executed.  Every call
# below is a 1:1 binding of a C-ABI function that IS exercised by the Python host (juliafem.jl_b200/_lib.py, api.py do
# the same walks) and by the GPU test-suite; argument order and types follow include/jfem_b200.h.
module JuliaFEMB200

using JuliaFEM
import JuliaFEM: initialize_backend, solve_backend!, assemble_elements!, GPU, Physics, Problem, Elasticity, Assembly, Element,
                 AbstractElasticityData, get_connectivity

const libjfem = get(ENV, "JFEM_B200_LIB", joinpath(@__DIR__, "..", "libjfem_b200.so"))

const JFEM_TET4, JFEM_HEX8, JFEM_TET10 = Cint(4), Cint(8), Cint(10)
const JFEM_TRI3, JFEM_QUAD4, JFEM_TRI6 = Cint(3), Cint(4), Cint(6)
const JFEM_MAT_LINEAR_ELASTIC, JFEM_MAT_NEO_HOOKEAN, JFEM_MAT_PERFECT_PLASTICITY, JFEM_MAT_STVK = Cint(0), Cint(1), Cint(2), Cint(3)
const JFEM_FIELD_STRAIN, JFEM_FIELD_STRESS = Cint(0), Cint(1)
const JFEM_PROJECT, JFEM_TANGENT, JFEM_USE_CSR, JFEM_JACOBI = Cint(1), Cint(2), Cint(4), Cint(8)

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
    remap::Dict{Int,Int}           # original node id -> 1..n
    n_dofs::Int
    f_ext::Vector{Float64}
    prescribed::Vector{Float64}
    fixed_dofs::Vector{Int64}
    nonlinear::Bool
    converged::Bool
    function ElasticityDataB200(handle, node_ids, remap)
        n = 3 * length(node_ids)
        d = new(handle, node_ids, remap, n, zeros(n), zeros(n), Int64[], false, true)
        finalizer(x -> (x.handle != C_NULL && ccall((:jfem_destroy, libjfem), Cint, (Ptr{Cvoid},), x.handle); x.handle = C_NULL), d)
        return d
    end
end

nnpe_code(::Type{JuliaFEM.Tet4}) = JFEM_TET4
nnpe_code(::Type{JuliaFEM.Hex8}) = JFEM_HEX8
nnpe_code(::Type{JuliaFEM.Tet10}) = JFEM_TET10
face_code(nn::Integer) = nn == 3 ? JFEM_TRI3 : nn == 4 ? JFEM_QUAD4 : nn == 6 ? JFEM_TRI6 :
    error("unsupported surface element with $nn nodes (supported: Tri3, Quad4, Tri6)")

"Flatten the body elements into the arrays jfem_create takes (same walk as initialize_gpu_data!, ext:86-132)."
function create_handle(elements::Vector, device::Integer)
    T = typeof(first(elements)).parameters[1]               # topology type parameter of Element{...}
    nn = Int(nnpe_code(T))
    conn_orig = [Int.(collect(get_connectivity(el))) for el in elements]
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

"Tuning knobs of the library (include/jfem_b200.h: \"patch_elems\", \"deterministic\", \"assembly_kernel\", \"geometric_stiffness\", ...)."
set_option!(handle::Ptr{Cvoid}, key::AbstractString, value::Real) =
    check(ccall((:jfem_set_option, libjfem), Cint, (Ptr{Cvoid}, Cstring, Cdouble), handle, key, Float64(value)))
set_option!(data::ElasticityDataB200, key::AbstractString, value::Real) = set_option!(data.handle, key, value)

"""
Material of the handle from per-element values: E_vec / nu_vec of ext:135-141.  A homogeneous mesh passes one parameter set,
otherwise the 2 x n_elems table goes down as per-element parameters.  `finite_strain` selects St. Venant-Kirchhoff -- Hooke's D
on the Green-Lagrange strain, which is what the classic path integrates (src/problems_elasticity.jl:255-332).
"""
function set_material!(handle::Ptr{Cvoid}, E::Vector{Float64}, nu::Vector{Float64}; finite_strain::Bool=false, geometric_stiffness::Bool=false)
    kind = finite_strain ? JFEM_MAT_STVK : JFEM_MAT_LINEAR_ELASTIC
    if all(==(E[1]), E) && all(==(nu[1]), nu)
        params = Float64[E[1], nu[1]]
        check(ccall((:jfem_set_material, libjfem), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Cint), handle, kind, params, 2, 0))
    else
        params = permutedims(hcat(E, nu))                    # 2 x n_elems, column-major: (E, nu) of element 1, then element 2, ...
        check(ccall((:jfem_set_material, libjfem), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Cint), handle, kind, params, 2, 1))
    end
    geometric_stiffness && set_option!(handle, "geometric_stiffness", 1)
    return nothing
end

"f (+)= consistent surface load of faces given by VOLUME-mesh node ids (jfem_surface_load; src/problems_elasticity.jl:454-502)."
function surface_load!(handle::Ptr{Cvoid}, f::Vector{Float64}, faces::Matrix{Int32}, traction::Union{Nothing,Matrix{Float64}},
                       pressure::Union{Nothing,Vector{Float64}})
    tp = traction === nothing ? Ptr{Float64}(C_NULL) : pointer(traction)
    pp = pressure === nothing ? Ptr{Float64}(C_NULL) : pointer(pressure)
    GC.@preserve traction pressure begin
        check(ccall((:jfem_surface_load, libjfem), Cint,
                    (Ptr{Cvoid}, Cint, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint),
                    handle, face_code(size(faces, 1)), size(faces, 2), faces, tp, pp, 1, f, 0))
    end
    return f
end

"f (+)= consistent body load, b = 3 x n_elems (\"displacement load\" of the volume elements, src/problems_elasticity.jl:412-426)."
body_load!(handle::Ptr{Cvoid}, f::Vector{Float64}, b::Matrix{Float64}) =
    (check(ccall((:jfem_body_load, libjfem), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Cint), handle, b, 1, 1, f, 0)); f)

function initialize_backend(backend::GPU, physics::Physics, time::Float64)
    elements = physics.body_elements
    handle, node_ids, remap = create_handle(elements, 0)
    data = ElasticityDataB200(handle, node_ids, remap)
    # per-element material, exactly the arrays the reference uploads (ext:135-141)
    E = Float64[el.fields.youngs_modulus for el in elements]
    nu = Float64[el.fields.poissons_ratio for el in elements]
    data.nonlinear = physics.properties.finite_strain
    set_material!(handle, E, nu; finite_strain=physics.properties.finite_strain,
                  geometric_stiffness=physics.properties.geometric_stiffness)
    # is_fixed / prescribed (ext:144-158): one entry of bc.node_ids per node, its components and their values
    bc = physics.bc_dirichlet
    vals = Float64[]
    for (i, node_id) in enumerate(bc.node_ids)
        haskey(remap, node_id) || continue
        for (comp_idx, comp) in enumerate(bc.components[i])
            push!(data.fixed_dofs, 3 * (remap[node_id] - 1) + comp)
            v = bc.values[i]
            push!(vals, v isa Number ? Float64(v) : Float64(v[comp_idx]))
        end
    end
    check(ccall((:jfem_set_dirichlet, libjfem), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Int64), handle, data.fixed_dofs, vals,
                length(data.fixed_dofs)))
    data.prescribed[data.fixed_dofs] .= vals
    # Neumann surfaces, integrated on the device.  For Tri3 the consistent one-point rule IS the lumped area/3 * t per node
    # of apply_surface_traction_kernel! (ext:368-416); Quad4 / Tri6 faces use the reference's GLQUAD4 / GLTRI3 rules.
    surf = physics.bc_neumann.surface_elements
    if !isempty(surf)
        by_type = Dict{Int,Vector{Int}}()
        for (i, s) in enumerate(surf)
            push!(get!(by_type, length(get_connectivity(s)), Int[]), i)
        end
        for (nn, idx) in by_type
            faces = zeros(Int32, nn, length(idx)); trac = zeros(Float64, 3, length(idx))
            for (q, i) in enumerate(idx)
                for (j, node) in enumerate(get_connectivity(surf[i]))
                    haskey(remap, Int(node)) || error("Surface element references node $node not in body mesh")   # ext:181
                    faces[j, q] = remap[Int(node)]
                end
                t = physics.bc_neumann.traction[i]
                trac[:, q] .= (t[1], t[2], t[3])
            end
            surface_load!(handle, data.f_ext, faces, trac, nothing)
        end
    end
    return data
end

function matvec(data::ElasticityDataB200, x::Vector{Float64}; flags::Cint=Cint(0))
    y = similar(x)
    check(ccall((:jfem_matvec, libjfem), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Cint), data.handle, x, y, flags, 0))
    return y
end

"Reaction forces of the constrained dofs, `la = f_int(u) - f_ext` there (the multipliers solve! returns, src/solvers.jl:205-216)."
function reactions(data::ElasticityDataB200, u::Vector{Float64})
    la = zeros(data.n_dofs)
    check(ccall((:jfem_reactions, libjfem), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint), data.handle, u, data.f_ext, la, 0))
    return la
end

"6 x n_nodes least-squares nodal fit of the Gauss-point strain / stress (postprocess!, src/problems_elasticity.jl:520-594)."
function nodal_recover(data::ElasticityDataB200, u::Vector{Float64}; field::Cint=JFEM_FIELD_STRESS)
    out = zeros(6, length(data.node_ids))
    check(ccall((:jfem_nodal_recover, libjfem), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Cint), data.handle, u, field, out, 0))
    return out
end

function solve_backend!(data::ElasticityDataB200, physics::Physics; tol=1e-6, max_iter=1000, newton_tol=1e-6, max_newton=20,
                        max_cg_per_newton=50)
    if data.nonlinear
        u = copy(data.prescribed)
        nit = Ref{Cint}(0); cgit = Ref{Cint}(0); res = Ref{Float64}(0.0)
        hist = zeros(Float64, 3, max_newton + 1)
        check(ccall((:jfem_newton_krylov, libjfem), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Cint, Cint, Float64, Float64, Cint, Ref{Cint}, Ref{Cint},
                     Ref{Float64}, Ptr{Float64}, Cint, Cint),
                    data.handle, data.f_ext, u, newton_tol, max_newton, max_cg_per_newton, 0.5, 0.9, 0, nit, cgit, res, hist,
                    max_newton + 1, 0))
        data.converged = res[] < newton_tol
        data.converged || @warn "Newton did not converge in $(nit[]) iterations: ||R|| = $(res[])"        # ext:850-853
        history = [(Int(hist[1, k]), hist[2, k], hist[3, k]) for k in 1:nit[]]
        return (u, Int(nit[]), Int(cgit[]), res[], history)
    end
    # lifting of prescribed values: f_I - K_IB u_B (src/solvers.jl:205-210); the operator is the pure K.v
    b = any(!iszero, data.prescribed) ? data.f_ext - matvec(data, data.prescribed) : copy(data.f_ext)
    x = zeros(data.n_dofs)
    it = Ref{Cint}(0); res = Ref{Float64}(0.0)
    check(ccall((:jfem_cg, libjfem), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Cint, Cint, Cint, Ref{Cint}, Ref{Float64}, Cint),
                data.handle, b, x, tol, 0, max_iter, 0, it, res, 0))
    data.converged = res[] < tol
    data.converged || @warn "CG did not converge in $(it[]) iterations: sqrt(r.r) = $(res[])"             # ext:630
    bI = copy(b); bI[data.fixed_dofs] .= 0.0
    return (x + data.prescribed, 1, Int(it[]), res[], [(Int(it[]), sqrt(sum(abs2, bI)), 0.0)])
end

"""
Classic path: `assemble!(problem::Problem{Elasticity}, time)` ends in `assemble_elements!` (src/assembly/assembly.jl:25).
This method integrates K(u) (and, for finite strain, f_int(u)) and the consistent body loads on the GPU and appends them to
`assembly.K` / `assembly.f` as COO triplets -- three bulk `append!`s, not one `add!` per entry -- so that `solve!(analysis)`
and everything downstream of the assembly work unchanged.  Surface elements (\"displacement traction force\", \"surface pressure\")
keep going through the reference's own `assemble!(..., Elasticity3DSurfaceElements)`: they are a separate element set there.
"""
function assemble_elements!(problem::Problem{Elasticity}, assembly::Assembly, elements::Vector{Element{T}}, time) where T
    handle, node_ids, remap = create_handle(elements, 0)
    props = problem.properties
    try
        E = Float64[el("youngs modulus", time) for el in elements]
        nu = Float64[el("poissons ratio", time) for el in elements]
        set_material!(handle, E, nu; finite_strain=props.finite_strain, geometric_stiffness=props.geometric_stiffness)
        n = Ref{Int64}(0); nnz = Ref{Int64}(0)
        check(ccall((:jfem_csr_size, libjfem), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), handle, n, nnz))
        rowptr = zeros(Int64, n[] + 1); colind = zeros(Int32, nnz[]); vals = zeros(Float64, nnz[])
        check(ccall((:jfem_csr_pattern, libjfem), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}), handle, rowptr, colind))
        # linearisation point: the current displacement of the elements (zero for the small-strain path)
        u = zeros(Float64, n[])
        if props.finite_strain
            for el in elements
                haskey(el, "displacement") || continue
                ue = el("displacement", time)
                for (k, node) in enumerate(get_connectivity(el)), c in 1:3
                    u[3 * (remap[Int(node)] - 1) + c] = ue[k][c]
                end
            end
        end
        fint = zeros(Float64, n[])
        check(ccall((:jfem_assemble_csr, libjfem), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint),
                    handle, props.finite_strain ? pointer(u) : Ptr{Float64}(C_NULL), vals, fint, 0, 0))
        # local dense numbering -> the problem's dofs 3*(node-1)+c (src/assembly/problems.jl:466-478)
        gdof = Vector{Int}(undef, n[])
        for (j, id) in enumerate(node_ids), c in 1:3
            gdof[3 * (j - 1) + c] = 3 * (id - 1) + c
        end
        rows = Vector{Int}(undef, nnz[])
        for r in 1:n[], p in rowptr[r]:(rowptr[r + 1] - 1)
            rows[p] = gdof[r]
        end
        append!(assembly.K.I, rows)
        append!(assembly.K.J, gdof[colind])
        append!(assembly.K.V, vals)
        # right-hand side: consistent body loads minus the internal force (f = f_ext - f_int, src/problems_elasticity.jl:407-426)
        b = zeros(Float64, 3, length(elements)); have_b = false
        for (e, el) in enumerate(elements), c in 1:3
            if haskey(el, "displacement load $c")
                b[c, e] = el("displacement load $c", time); have_b = true
            end
        end
        f = zeros(Float64, n[])
        have_b && body_load!(handle, f, b)
        f .-= fint
        nz = findall(!iszero, f)
        append!(assembly.f.I, gdof[nz]); append!(assembly.f.J, ones(Int, length(nz))); append!(assembly.f.V, f[nz])
    finally
        ccall((:jfem_destroy, libjfem), Cint, (Ptr{Cvoid},), handle)
    end
    return nothing
end

end # module
