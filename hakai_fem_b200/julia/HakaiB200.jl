# HakaiB200.jl — thin ccall binding of libhakai_b200.so (include/hakai_b200.h) for HAKAI_j.jl.
#
# NOT EXECUTED IN THE BUILD IMAGE (no Julia there); it is the binding a maintainer adds next to HAKAI_j.jl.
# Every wrapper passes the reference's own arrays (column-major, 1-based Int64 / Float64) untouched: the
# library converts layouts itself.  INTEGRATION.md shows the patch to hakai().
module HakaiB200

const LIB = get(ENV, "HAKAI_B200_LIB", joinpath(@__DIR__, "..", "libhakai_b200.so"))

# struct hk_params (include/hakai_b200.h); field order and types must match
mutable struct Params
    struct_size::Int32
    device::Int32
    d_time::Float64
    element_min_size::Float64
    element_max_size::Float64
    contact_flag::Int32
    triax_route::Int32
    contact_d_lim_factor::Float64
    contact_myu::Float64
    contact_kc_other::Float64
    contact_kc_self::Float64
    contact_cr_other::Float64
    contact_cr_self::Float64
    contact_ddiv_other::Float64
    contact_ddiv_self::Float64
    deterministic::Int32
    element_mode::Int32           # 0 fast element kernel, 1 reference-order kernel (bit-identical to the CPU oracle)
    contact_dmax_clamp::Int32     # 1: v0.0.1's penetration-rate clamp (HAKAI-v0.0.1 HAKAI_j.jl:2756)
    reserved0::Int32
    Params() = new()
end

struct Engine
    ptr::Ptr{Cvoid}
end

function check(e, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:hk_last_error, LIB), Cstring, (Ptr{Cvoid},), e))
    error("libhakai_b200: code $rc: $msg")
end

function default_params()
    p = Params()
    rc = ccall((:hk_default_params, LIB), Cint, (Ref{Params},), p)
    rc == 0 || error("hk_default_params failed")
    return p
end

function create(p::Params)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:hk_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{Params}), out, p)
    check(C_NULL, rc)          # fails loudly when there is no CUDA device: no CPU fallback
    return Engine(out[])
end

destroy(e::Engine) = ccall((:hk_destroy, LIB), Cint, (Ptr{Cvoid},), e.ptr)

function csr(lists)
    ptr = Int64[0]
    flat = Int64[]
    for l in lists
        append!(flat, l)
        push!(ptr, length(flat))
    end
    return ptr, flat
end

set_mesh(e, coordmat::Matrix{Float64}, elementmat::Matrix{Int}, element_material::Vector{Int},
         element_instance::Vector{Int}, diag_M::Vector{Float64}) =
    check(e.ptr, ccall((:hk_set_mesh, LIB), Cint,
          (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
          e.ptr, size(coordmat, 2), size(elementmat, 2), coordmat, elementmat, element_material, element_instance, diag_M))

function add_material(e, m)    # m::MaterialType (readInpFile_j.jl:84-96)
    npp = size(m.plastic, 1); nd = size(m.ductile, 1)
    check(e.ptr, ccall((:hk_add_material, LIB), Cint,
          (Ptr{Cvoid}, Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
          e.ptr, m.young, m.poisson, m.density, npp, npp > 0 ? m.plastic : C_NULL, npp > 1 ? m.Hd : C_NULL,
          nd, nd > 0 ? m.ductile : C_NULL))
end

function add_bc(e, bc)         # bc::BCType (readInpFile_j.jl:98-104)
    ptr, flat = csr(bc.dof)
    has_amp = length(bc.amp_name) > 0
    at = has_amp ? Float64.(bc.amplitude.time) : Float64[]
    av = has_amp ? Float64.(bc.amplitude.value) : Float64[]
    check(e.ptr, ccall((:hk_add_bc, LIB), Cint,
          (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
          e.ptr, length(bc.dof), ptr, flat, Float64.(bc.value), length(at), at, av))
end

function add_ic(e, ic)         # ic::ICType (readInpFile_j.jl:106-111)
    ptr, flat = csr(ic.dof)
    check(e.ptr, ccall((:hk_add_ic, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
          e.ptr, length(ic.dof), ptr, flat, Float64.(ic.value)))
end

add_instance(e, inst) =        # inst::InstanceType after get_element_face (HAKAI_j.jl:255-261)
    check(e.ptr, ccall((:hk_add_instance, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}),
          e.ptr, inst.node_offset, inst.nNode, inst.element_offset, inst.nElement, inst.surfaces, inst.surfaces_eleid))

add_contact_pair(e, i_instance, j_instance, ct) =   # ct::ContactTriangle (HAKAI_j.jl:72-78, 361-398)
    check(e.ptr, ccall((:hk_add_contact_pair, LIB), Cint,
          (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Float64),
          e.ptr, i_instance, j_instance, length(ct.c_nodes_i), ct.c_nodes_i, length(ct.c_nodes_j), ct.c_nodes_j,
          size(ct.c_triangles, 1), ct.c_triangles, ct.c_triangles_eleid, ct.young))

# Contact set-up on the GPU instead of get_element_face / get_surface_triangle / the CT loop (HAKAI_j.jl:250-398):
# replaces the add_instance / add_contact_pair calls.  INSTANCE, MATERIAL, CP = MODEL's arrays (CP empty: ALL EXTERIOR).
function build_contact(e, INSTANCE, MATERIAL, CP)
    no = Int64[i.node_offset for i in INSTANCE]; nn = Int64[i.nNode for i in INSTANCE]
    eo = Int64[i.element_offset for i in INSTANCE]; ne = Int64[i.nElement for i in INSTANCE]
    yg = Float64[MATERIAL[i.material_id].young for i in INSTANCE]
    i1 = Int64[c.instance_id_1 for c in CP]; i2 = Int64[c.instance_id_2 for c in CP]
    p1, e1 = csr([c.elements_1 for c in CP]); p2, e2 = csr([c.elements_2 for c in CP])
    check(e.ptr, ccall((:hk_build_contact, LIB), Cint,
          (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64},
           Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
          e.ptr, length(INSTANCE), no, nn, eo, ne, yg, length(CP), i1, i2, p1, e1, p2, e2))
end

finalize!(e) = check(e.ptr, ccall((:hk_finalize, LIB), Cint, (Ptr{Cvoid},), e.ptr))

function step!(e, t_first::Integer, n_steps::Integer)
    nd = Ref{Int64}(0)
    check(e.ptr, ccall((:hk_step, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ref{Int64}), e.ptr, t_first, n_steps, nd))
    return nd[]
end

download!(e, disp, velo, integ_stress, integ_strain, integ_eq_plastic_strain, integ_triax_stress, element_flag) =
    check(e.ptr, ccall((:hk_download, LIB), Cint,
          (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
          e.ptr, disp, velo, integ_stress, integ_strain, integ_eq_plastic_strain, integ_triax_stress, element_flag))

# cal_node_stress_strain (HAKAI_j.jl:3408-3486) on the device: fills the NodeDataType arrays (node_stress and node_strain
# are (nNode,6) Julia matrices).  Replaces `node_data = cal_node_stress_strain(...)` at HAKAI_j.jl:478, 936.
node_output!(e, nd) =          # nd::NodeDataType (HAKAI_j.jl:43-50)
    check(e.ptr, ccall((:hk_node_output, LIB), Cint,
          (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32),
          e.ptr, nd.node_stress, nd.node_strain, nd.node_eq_plastic_strain, nd.node_mises_stress, nd.node_triax_stress,
          C_NULL, 0))

# state summary reduced on the device: (live elements, min / max eq. plastic strain of live Gauss points, yielded points)
function state_summary(e)
    out = zeros(Float64, 8)
    check(e.ptr, ccall((:hk_state_summary, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.ptr, out))
    return (live_elements = Int(out[1]), eps_min = out[2], eps_max = out[3], yielded_points = Int(out[4]))
end

function deleted_ids(e)
    n = Ref{Int64}(0)
    check(e.ptr, ccall((:hk_deleted_ids, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ref{Int64}), e.ptr, C_NULL, 0, n))
    ids = zeros(Int64, n[]); steps = zeros(Int64, n[])
    check(e.ptr, ccall((:hk_deleted_ids, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ref{Int64}), e.ptr, ids, n[], n))
    check(e.ptr, ccall((:hk_deleted_steps, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ref{Int64}), e.ptr, steps, n[], n))
    return ids, steps
end

# ---- multi-GPU: one Julia process (or task) per GPU, element-block partition ------------------------------------------
# The engine owns the NCCL communicator, so the Julia side needs NO NCCL binding: rank 0 asks the library for the
# 128-byte id, the host hands it to the other ranks by whatever it already has (MPI.Bcast!, a file, a socket), and from
# then on `step!(e, t, n)` runs n complete multi-GPU steps (pack -> ncclSend/ncclRecv -> split step) inside the library.
#   halo_nodes[i]: local 1-based ids of the nodes shared with neighbour i, ascending global id (before finalize!)
#   ranks[i]:      global rank of neighbour i
set_halo(e, halo_nodes::Vector{Vector{Int}}) = begin
    ptr, flat = csr(halo_nodes)
    check(e.ptr, ccall((:hk_set_halo, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}), e.ptr, length(halo_nodes), ptr, flat))
end
set_halo_ranks(e, my_rank::Integer, ranks::Vector{Int}) =
    check(e.ptr, ccall((:hk_set_halo_ranks, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}), e.ptr, my_rank, length(ranks), ranks))
function comm_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:hk_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)
    rc == 0 || check(C_NULL, rc)
    return id
end
comm_init(e, id::Vector{UInt8}, rank::Integer, world::Integer) =
    check(e.ptr, ccall((:hk_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), e.ptr, id, rank, world))

# contact across ranks (INTEGRATION.md "Decks with contact"): node lists 0 own surface nodes / 1 ghost copies / 2 all
# surface nodes (1-based local ids), then the engine exchanges by itself; with failure: global maps + device-side replay
set_node_list(e, which::Integer, nodes::Vector{Int64}) =
    check(e.ptr, ccall((:hk_set_node_list, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}), e.ptr, which, length(nodes), nodes))
comm_contact(e, maxlen::Integer, src_index::Vector{Int64}) =
    check(e.ptr, ccall((:hk_comm_contact, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}), e.ptr, maxlen, src_index))
set_global_maps(e, node_map::Vector{Int64}, elem_map::Vector{Int64}, element_instance::Vector{Int64}) =
    check(e.ptr, ccall((:hk_set_global_maps, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}),
                       e.ptr, length(node_map), node_map, length(elem_map), elem_map, element_instance))
comm_erosion(e, max_deleted_per_step::Integer = 4096) =
    check(e.ptr, ccall((:hk_comm_erosion, LIB), Cint, (Ptr{Cvoid}, Int32), e.ptr, max_deleted_per_step))

end # module
