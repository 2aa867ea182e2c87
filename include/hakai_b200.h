/*
 * hakai_b200.h — C ABI of the B200-native HAKAI time-step engine (libhakai_b200.so).
 *
 * The reference (yozoyugen/HAKAI-fem) has no FFI: its per-step hot path is inline Julia in
 * hakai() (HAKAI-v0.0.2/Julia/HAKAI_j.jl:487-951, cited below as J2:<line>).  This header is the
 * boundary a maintainer binds with `ccall` to replace that loop body; INTEGRATION.md shows the
 * Julia stub.  Conventions:
 *   - plain C, no exceptions; every call returns 0 on success or a negative hk_status code,
 *     with a message available from hk_last_error().
 *   - all pointers are HOST pointers owned by the caller, in the reference's own (Julia,
 *     column-major, 1-based, Int64/Float64) layouts; the library copies and never keeps them.
 *   - the engine owns all device memory; one host thread drives one engine; not re-entrant.
 *   - fn = 3*nNode, nip = 8*nElement.  Voigt order xx,yy,zz,xy,yz,xz (engineering shear).
 *
 * The same signatures with prefix `hko_` are exported by oracle/libhakai_oracle.so (the CPU
 * restatement used ONLY by tests/ and bench.py's cpu_baseline) so the parity tests can drive
 * both through one wrapper.
 */
#ifndef HAKAI_B200_H
#define HAKAI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hk_engine hk_engine;

enum hk_status {
    HK_OK = 0,
    HK_ERR_ARG = -1,        /* bad argument / call order                      */
    HK_ERR_CUDA = -2,       /* CUDA runtime error (message has the detail)    */
    HK_ERR_NO_DEVICE = -3,  /* no sm_100 device: there is NO CPU fallback     */
    HK_ERR_STATE = -4,      /* engine not finalised / already finalised       */
    HK_ERR_UNSUPPORTED = -5
};

/* Every constant the reference hard-codes in source (SURVEY §5 "Config / flags").
 * hk_default_params() fills the reference's values. */
typedef struct hk_params {
    int32_t struct_size;        /* = sizeof(hk_params), ABI check                               */
    int32_t device;             /* CUDA device ordinal                                          */
    double  d_time;             /* MODEL.d_time*sqrt(mass_scaling)                J2:114        */
    double  element_min_size;   /* elementMinSize                                 J2:418        */
    double  element_max_size;   /* elementMaxSize                                 J2:420        */
    int32_t contact_flag;       /* MODEL.contact_flag (0 none, 1 contact, 2 +self) J2:100       */
    int32_t triax_route;        /* 0 = invariants (I1/3)/sqrt(3 J2); 1 = closed-form 3x3 eigen-
                                   values as StaticArrays.eigvals (oracle only)    J2:1004-1016 */
    double  contact_d_lim_factor; /* d_lim = factor*elementMinSize, 0.3           J2:2254       */
    double  contact_myu;          /* friction coefficient, 0.25                   J2:2255       */
    double  contact_kc_other;     /* kc_o = 1                                     J2:2256       */
    double  contact_kc_self;      /* kc_s = 1                                     J2:2257       */
    double  contact_cr_other;     /* Cr_o = 0                                     J2:2258       */
    double  contact_cr_self;      /* Cr_s = 0                                     J2:2259       */
    double  contact_ddiv_other;   /* cell size factor 1.1*elementMaxSize          J2:2331       */
    double  contact_ddiv_self;    /* 0.6*elementMaxSize for self contact          J2:2333       */
    int32_t deterministic;        /* 1 (default): reproducible assembly and contact sums        */
    int32_t element_mode;         /* 0 (default): fast element kernel (mode form, FMA) — within 1e-13/step of the
                                     reference arithmetic.  1: reference-order kernel: the dense 6x24 Bfinal algebra of
                                     J2:1033-1371 in the reference's operation order without FMA — bit-identical to the
                                     CPU oracle (and ~10x slower); resolves contact ties like the reference          */
    int32_t contact_dmax_clamp;   /* 0 (default, v0.0.2 behaviour): off.  1: v0.0.1's penetration-rate clamp — a slave node's
                                     penetration may grow by at most d_max = max_n |d_disp_n| of the previous step:
                                     if d - d_node_pre[i] > d_max: d = d_node_pre[i] + d_max
                                     (HAKAI-v0.0.1/Julia/HAKAI_j.jl:2756, 3218; d_node J1:2898, d_max J1:618).  Single-domain
                                     engines only                                                                       */
    int32_t reserved0;            /* keep 0 */
} hk_params;

int hk_default_params(hk_params* p);

/* Allocates the engine on p->device.  Fails with HK_ERR_NO_DEVICE when no CUDA device exists. */
int hk_create(hk_engine** out, const hk_params* p);
int hk_destroy(hk_engine* e);
const char* hk_last_error(const hk_engine* e);   /* e may be NULL: last create() error */

/* Mesh (replaces the locals of hakai(), J2:91-98, 186-218).
 *   coordmat          f64 (3,nNode) column-major                      J2:93
 *   elementmat        i64 (8,nElement) column-major, 1-based          J2:96
 *   element_material  i64 (nElement) 1-based material index            J2:97
 *   element_instance  i64 (nElement) 1-based instance index            J2:98 (may be NULL: all 1)
 *   diag_M            f64 (fn); the 3 dofs of a node must hold the same mass (J2:205-215)
 */
int hk_set_mesh(hk_engine* e, int64_t nNode, int64_t nElement,
                const double* coordmat, const int64_t* elementmat,
                const int64_t* element_material, const int64_t* element_instance,
                const double* diag_M);

/* One call per MODEL.MATERIAL entry, in order (MaterialType, readInpFile_j.jl:84-96).
 *   plastic  f64 (npp,2) column-major [yield stress | eq. plastic strain]; npp = 0: elastic
 *   Hd       f64 (npp-1) hardening slopes                readInpFile_j.jl:763-768
 *   ductile  f64 (nd,3) column-major [eps_f | triax | rate]; nd = 0: no deletion
 * Dmat and G are rebuilt from young/poisson exactly as J2:143-159. */
int hk_add_material(hk_engine* e, double young, double poisson, double density,
                    int64_t npp, const double* plastic, const double* Hd,
                    int64_t nd, const double* ductile);

/* One call per MODEL.BC entry, in order (later entries override earlier ones, J2:585-617).
 *   dof_ptr  i64 (n_lists+1) CSR offsets (0-based) into dofs
 *   dofs     i64 1-based dof ids (BC[i].dof[j])
 *   values   f64 (n_lists)   BC[i].value[j]
 *   n_amp = 0: amp = 1; else piecewise-linear table amp_time/amp_value (n_amp >= 2). */
int hk_add_bc(hk_engine* e, int64_t n_lists, const int64_t* dof_ptr, const int64_t* dofs,
              const double* values, int64_t n_amp, const double* amp_time, const double* amp_value);

/* Initial velocity (J2:233-239): disp_pre[dof] = -value*d_time, velo[dof] = value.
 * Call once per MODEL.IC entry, in order. */
int hk_add_ic(hk_engine* e, int64_t n_lists, const int64_t* dof_ptr, const int64_t* dofs,
              const double* values);

/* Contact set-up (J2:250-398).  One hk_add_instance per MODEL.INSTANCE (in order) with the
 * outward-oriented element faces of get_element_face (J2:1946-1992), part-local 1-based node
 * ids; then one hk_add_contact_pair per ContactTriangle CT[c] (J2:357-398), global 1-based ids. */
int hk_add_instance(hk_engine* e, int64_t node_offset, int64_t nNode,
                    int64_t element_offset, int64_t nElement,
                    const int64_t* surfaces /* (6*nElement,4) col-major */,
                    const int64_t* surfaces_eleid /* (6*nElement) */);
int hk_add_contact_pair(hk_engine* e, int64_t i_instance, int64_t j_instance,
                        int64_t nn_i, const int64_t* c_nodes_i,
                        int64_t nn_j, const int64_t* c_nodes_j,
                        int64_t nTri, const int64_t* c_triangles /* (nTri,3) col-major */,
                        const int64_t* c_triangles_eleid, double young);

/* Contact set-up on the device, instead of hk_add_instance + hk_add_contact_pair (A12: get_element_face J2:1946-1992,
 * get_surface_triangle J2:1996-2164, pair list J2:272-398).  The faces of every instance are oriented, sorted by their
 * node set and matched on the GPU (radix sort, O(F log F); the reference's face matching is O(F^2)), with the
 * reference's quirks (odd groups emit their last member; the very last face is never emitted, J2:2040); the face each
 * deletion would expose (add_surface_triangle, J2:2167-2245) comes out of the same sort.  Call after hk_set_mesh.
 *   instances: node_offset / nNode / element_offset / nElement as InstanceType (readInpFile_j.jl:59-76), young[i] =
 *              MATERIAL[INSTANCE[i].material_id].young (J2:372)
 *   n_cp = 0:  *Contact Inclusions, ALL EXTERIOR — pairs as J2:272-314 (params.contact_flag == 2 adds the self pairs)
 *   n_cp > 0:  MODEL.CP: instance ids (1-based) of both sides and, per side, the *Surface element set as CSR lists of
 *              part-local 1-based element ids (cp_ptrX / cp_elemsX may be NULL or a list empty: all elements)
 * Orientation uses the global undeformed coordinates (the reference uses the part's: identical up to the rigid
 * instance transform). */
int hk_build_contact(hk_engine* e, int64_t n_inst, const int64_t* node_offset, const int64_t* nNode,
                     const int64_t* element_offset, const int64_t* nElement, const double* young, int64_t n_cp,
                     const int64_t* cp_inst1, const int64_t* cp_inst2, const int64_t* cp_ptr1, const int64_t* cp_elems1,
                     const int64_t* cp_ptr2, const int64_t* cp_elems2);

/* Freezes the set-up: builds device layouts (SoA state, node->element gather table, BC table,
 * contact buckets), initialises state as J2:220-230, 447-465 (zero state, yield = plastic[1,1],
 * element_flag = 1).  Must be called once before hk_step/hk_upload_state/hk_download. */
int hk_finalize(hk_engine* e);

/* Runs steps t = t_first .. t_first+n_steps-1 of the loop J2:487-951 (t is 1-based;
 * current_time = t*d_time, J2:589).  Synchronous on return.  *n_deleted_out (may be NULL)
 * receives the number of elements deleted during these steps (J2:733-736). */
int hk_step(hk_engine* e, int64_t t_first, int64_t n_steps, int64_t* n_deleted_out);

/* hk_step = hk_step_enqueue + hk_sync.  hk_step_enqueue only enqueues the kernels on the engine's stream (it
 * still synchronises internally when contact surfaces may change after a deletion); hk_sync waits for them and
 * reports the elements deleted since the previous hk_sync/hk_step.  The multi-GPU driver enqueues
 * pack -> send/recv -> step for many steps without blocking the host. */
int hk_step_enqueue(hk_engine* e, int64_t t_first, int64_t n_steps);
int hk_sync(hk_engine* e, int64_t* n_deleted_out);

/* Output taps (A13): fills caller-owned arrays in the reference's layouts; NULL = skip.
 *   disp, velo                       f64 (fn)
 *   integ_stress, integ_strain       f64 (6,nip) column-major
 *   integ_eq_plastic_strain, integ_triax_stress   f64 (nip)
 *   element_flag                     i64 (nElement) */
int hk_download(hk_engine* e, double* disp, double* velo,
                double* integ_stress, double* integ_strain,
                double* integ_eq_plastic_strain, double* integ_triax_stress,
                int64_t* element_flag);

/* Remaining loop-carried state, for tests / restart.  NULL = skip.
 *   disp_pre, Q, external_force f64 (fn); position f64 (3,nNode); integ_yield_stress f64 (nip);
 *   elementVolume f64 (nElement)  (J2:1169) */
int hk_download_ex(hk_engine* e, double* disp_pre, double* Q, double* external_force,
                   double* position, double* integ_yield_stress, double* elementVolume);

/* Output tap computed on the device: cal_node_stress_strain (J2:3408-3486).  Gauss points -> element mean ->
 * mean over the elements incident to a node (deleted ones included, in ascending element order like the reference's
 * loop) -> von Mises of the nodal stress.  Fills NodeDataType (J2:43-50) arrays in Julia's layout; NULL = skip.
 *   node_stress, node_strain f64 (nNode,6) column-major; node_eq_plastic_strain, node_mises_stress,
 *   node_triax_stress f64 (nNode); inc_num f64 (nNode) = number of incident elements (J2:3456-3460).
 * raw != 0: sums are NOT divided by inc_num and node_mises_stress is not written — for partitioned meshes, where
 * the host adds the neighbours' sums of the interface nodes first.  Moves 15 doubles per node to the host instead
 * of 112 per element.  The first call allocates 14 doubles per element + 16 per node of device work space, which the
 * engine keeps until hk_destroy. */
int hk_node_output(hk_engine* e, double* node_stress, double* node_strain, double* node_eq_plastic_strain,
                   double* node_mises_stress, double* node_triax_stress, double* inc_num, int32_t raw);

/* Overwrite loop-carried state (tests, restart).  NULL = keep.  Layouts as hk_download*. */
int hk_upload_state(hk_engine* e, const double* disp, const double* disp_pre, const double* velo,
                    const double* Q, const double* integ_stress, const double* integ_strain,
                    const double* integ_eq_plastic_strain, const double* integ_yield_stress,
                    const int64_t* element_flag);

/* Ids (1-based, deletion order: ascending step, ascending id within a step, J2:701-735) of
 * all elements deleted since hk_finalize.  Copies min(cap, n) ids; *n_out = total count. */
int hk_deleted_ids(hk_engine* e, int64_t* ids, int64_t cap, int64_t* n_out);

/* The step t (as passed to hk_step) in which each of those elements was deleted, aligned with hk_deleted_ids; 0 for
 * ids replayed through hk_apply_deleted.  Lets a driver that enqueued many steps in one call reconstruct the
 * live-element count of every step without synchronising in between. */
int hk_deleted_steps(hk_engine* e, int64_t* steps, int64_t cap, int64_t* n_out);

/* Current contact surface of pair c (0-based), after A10 updates (J2:767-804): sizes, or
 * arrays when non-NULL (caller sizes them from a first call). */
int hk_contact_pair_info(hk_engine* e, int64_t c, int64_t* nn_i, int64_t* nn_j, int64_t* nTri,
                         int64_t* c_nodes_i, int64_t* c_nodes_j,
                         int64_t* c_triangles, int64_t* c_triangles_eleid);

/* Counters since finalize: [0] negative-Jacobian warnings (J2:1736-1739), [1] contact hits,
 * [2] contact candidate tests, [3] kernel launches, [4] steps run. */
int hk_counters(hk_engine* e, int64_t out[8]);

/* Cheap summary of the element state, reduced on the device (32 bytes cross the bus): what a driver needs to
 * know the regime it is timing and how many elements are still alive, without downloading nip doubles.
 *   out[0] live elements (element_flag == 1, J2:733); out[1] / out[2] min / max of integ_eq_plastic_strain over the
 *   Gauss points of live elements; out[3] number of those Gauss points with eps > 0; out[4..7] = 0 (reserved). */
int hk_state_summary(hk_engine* e, double out[8]);

/* Per-kernel device timing with CUDA events on the engine's stream.
 *   kind: 0 contact, 1 nodal update (+assembly gather, BC, kinematics), 2 element, 3 other.
 * hk_profile(e, 1) enables/reset; hk_profile_read returns total ms and launch count per kind. */
int hk_profile(hk_engine* e, int32_t enable);
int hk_profile_read(hk_engine* e, double ms[4], int64_t launches[4]);
/* The same with the kinds a multi-GPU / fracture run adds: 4 halo exchange (the ncclSend/ncclRecv group of the engine's
 * own communicator, timed on its side stream — it overlaps kind 1), 5 deletion pass (ordered list, zeroing, exposed
 * faces), 6 time between the end of one profiled launch and the start of the next on the engine's stream (launch gaps;
 * its count is the number of gaps); 7 reserved. */
int hk_profile_read_ex(hk_engine* e, double ms[8], int64_t launches[8]);

/* Run all work on the caller's CUDA stream (cudaStream_t as void*).  NULL is the CUDA legacy default stream (what
 * torch.cuda.current_stream().cuda_stream returns by default), so that NCCL ops enqueued by the host framework are
 * ordered with the engine's kernels.  Without this call the engine uses a private non-blocking stream. */
int hk_set_stream(hk_engine* e, void* cuda_stream);

/* ---- multi-GPU: one engine per rank over an element-block partition ------------------------
 * Nodes on a partition interface exist on both ranks; each rank computes the internal force of its own
 * elements only, so before the nodal update the partial sums must be exchanged and added.  The engine
 * packs / unpacks; the transport (ncclSend/ncclRecv, here through torch.distributed) is the host's job:
 *     hk_halo_pack(e)  ->  exchange send/recv buffers with every neighbour  ->  hk_step(e, t, 1, ..)
 * Both ranks then update the shared nodes redundantly from identical inputs (a+b == b+a bitwise), so no
 * position exchange is needed.
 *   hk_set_halo       before hk_finalize.  nodes: LOCAL 1-based node ids shared with neighbour i, in an order
 *                     both sides agree on (ascending global id); nbr_ptr: CSR offsets (n_neighbors+1).
 *   hk_halo_bind      after hk_finalize.  send/recv: DEVICE buffers of 3*n_nodes(i) doubles owned by the caller
 *                     (e.g. torch tensors registered with NCCL).
 *   hk_halo_pack      fills every send buffer with this rank's partial Q of the shared nodes (stream-ordered).
 * hk_step() adds the received partials, summed in neighbour order, to the gathered internal force. */
int hk_set_halo(hk_engine* e, int64_t n_neighbors, const int64_t* nbr_ptr, const int64_t* nodes);
int hk_halo_bind(hk_engine* e, int64_t neighbor, void* send_dev, void* recv_dev);
int hk_halo_pack(hk_engine* e);
/* Global rank of this engine and of every hk_set_halo neighbour (same order).  With them the partial forces of an
 * interface node are summed one holder at a time in ascending rank order, this rank's own partial in its place, so that
 * EVERY holder forms the same bits even when a node is shared by three or more ranks; without them the own partial comes
 * first (identical on both sides only for nodes with exactly two holders). */
int hk_set_halo_ranks(hk_engine* e, int64_t my_rank, int64_t n_neighbors, const int64_t* ranks);

/* NCCL inside the library: the engine owns its communicator (libnccl.so.2 is loaded at run time; the copy already in
 * the process is shared when there is one).
 *   hk_comm_unique_id  fills 128 bytes (an ncclUniqueId) on one rank; the host hands them to every rank by any means.
 *   hk_comm_init       after hk_finalize and hk_set_halo_ranks: creates the communicator on the engine's device and
 *                      allocates the engine's own send/recv blocks (hk_halo_bind is then not needed).
 * With a communicator hk_step / hk_step_enqueue(t, n) run n >= 1 complete multi-GPU steps with no host involvement:
 * pack -> ncclSend/ncclRecv with every neighbour on a side stream, overlapped with the nodal update of the non-interface
 * nodes -> interface nodes -> element kernel. */
int hk_comm_unique_id(void* id128);
int hk_comm_init(hk_engine* e, const void* id128, int32_t rank, int32_t world);
/* Contact across ranks exchanged by the engine too.  After hk_set_node_list(0 / 1 / 2) (own surface nodes, ghost copies,
 * all surface nodes — see "multi-GPU contact" below): maxlen = the longest own-export list of any rank (every rank's
 * block of the all-gather is padded to it), src_index[i] = rank * maxlen + position of ghost i's owner record.  From
 * then on every step of hk_step / hk_step_enqueue first runs, on the engine's stream: export of the own surface nodes ->
 * ncclAllGather -> ghost copies -> contact pass on the local master triangles -> 43-bit limbs of the 128-bit force
 * accumulators -> ncclAllReduce(int64, sum) (exact) -> the halo step.  Call again whenever the lists change.  Surfaces
 * that erode across ranks (hk_set_global_maps) need the host's replay after every step (then n_steps = 1) unless
 * hk_comm_erosion moved that replay onto the device. */
int hk_comm_contact(hk_engine* e, int64_t maxlen, const int64_t* src_index);

/* The asynchronous step calls imply no output frame, so they do not store integ_triax_stress (hk_download and
 * hk_node_output then derive it from the current stress).  hk_mark_frame announces that the last step of the NEXT
 * hk_step_enqueue / hk_step_finish call is followed by a frame: that step stores the triaxiality computed inside it
 * (J2:677) — for elements deleted in that very step this is the value from before their stress is zeroed, as in
 * a frame written by the reference. */
int hk_mark_frame(hk_engine* e);

/* Split form of hk_step_enqueue(e, t, 1) that overlaps the exchange with compute:
 *     hk_halo_pack -> start send/recv -> hk_step_begin(t) [contact, nodal update of all non-interface nodes]
 *                  -> wait send/recv  -> hk_step_finish(t) [add partials, interface nodes, element kernel] */
int hk_step_begin(hk_engine* e, int64_t t);
int hk_step_finish(hk_engine* e, int64_t t);

/* ---- multi-GPU contact: contact-surface nodes all-gathered ---------------------------------------
 * Every rank holds the GLOBAL contact node lists (nodes it does not own are appended to its mesh as "ghost" nodes
 * that no element references) and the master triangles of the elements it owns.  Per step, before hk_step_*:
 *     hk_nodes_export(own)   -> all-gather {position, velocity} of the surface nodes each rank owns
 *     hk_nodes_import(ghost) -> ghost copies refreshed
 *     hk_contact_enqueue     -> bounding boxes / cells / narrow phase on the local triangles (same cell grid and
 *                               hit tests as single-GPU because the node lists are global)
 *     hk_contact_export      -> all-gather the 128-bit fixed-point force accumulators of the surface nodes
 *     hk_contact_import      -> exact integer sum over ranks (order independent => identical on every rank)
 * then the usual (halo) step, which skips its own contact pass once.
 *   list ids: LOCAL 1-based node ids; buffers: DEVICE memory owned by the caller.
 *   export/import node record: 6 doubles {x,y,z,vx,vy,vz}; accumulator record: 6 x uint64 per node. */
int hk_set_node_list(hk_engine* e, int32_t which /* 0 own-export, 1 ghost-import, 2 surface (force exchange),
                                                    3 state-export, 4 state-import (ghost-element partitions) */,
                     int64_t n, const int64_t* nodes);
int hk_nodes_export(hk_engine* e, void* out_dev);                 /* list 0 -> 6 doubles per node            */
int hk_nodes_import(hk_engine* e, const void* in_dev, const int64_t* src_index /* host, n(list 1) */);
int hk_contact_enqueue(hk_engine* e);                             /* contact pass of the NEXT step, now     */
int hk_contact_export(hk_engine* e, void* out_dev);               /* list 2 -> 6 uint64 per node             */
int hk_contact_import(hk_engine* e, const void* in_dev, int64_t n_ranks);   /* sum of n_ranks records per node */
/* Ghost-element partitions (optional mode): a rank holds its element block PLUS every element that shares
 * a node with it, so all nodes of its own elements see their complete force sum locally, in the global ascending
 * element order — results are bit-identical for any number of ranks and no force halo is needed.  What is
 * exchanged instead is the displacement state of the outer nodes of the ghost layer, once per step between the nodal
 * update and the element kernel:   hk_step_begin(t) -> hk_state_export (list 3: 6 doubles {disp, disp_pre} per node)
 * -> send/recv -> hk_state_import (list 4, same record order) -> hk_step_finish(t).   No hk_set_halo in this mode. */
int hk_state_export(hk_engine* e, void* out_dev);
int hk_state_import(hk_engine* e, const void* in_dev);

/* The same exchange for a plain integer all-reduce (ncclAllReduce, ncclInt64, ncclSum) instead of an all-gather: every
 * 128-bit accumulator travels as three 43-bit limbs in int64 lanes — 9 x int64 per surface node, summed lane-wise by
 * the collective, recombined exactly (mod 2^128) on import; valid for fewer than 2^20 ranks. */
int hk_contact_export_limbs(hk_engine* e, void* out_dev);      /* 9 x i64 per surface node */
int hk_contact_import_limbs(hk_engine* e, const void* in_dev); /* lane-wise sums over all ranks */
/* Erosion of contact surfaces across ranks (add_surface_triangle, J2:767-804, 2167-2245).  The instance face tables
 * given to hk_add_instance are then the GLOBAL ones; the maps translate global 1-based ids to this rank's local
 * 1-based ids (0 = not present / not owned).  After every step the host all-gathers the freshly deleted GLOBAL
 * element ids (ascending) and every rank applies the same list: node lists grow identically everywhere, each new
 * triangle is kept by the rank that owns its element. */
int hk_set_global_maps(hk_engine* e, int64_t n_global_nodes, const int64_t* node_map,
                       int64_t n_global_elements, const int64_t* elem_map, const int64_t* element_instance);
int hk_apply_deleted(hk_engine* e, int64_t n, const int64_t* global_ids);
/* The same replay ON THE DEVICE (call after hk_set_global_maps, before the first step).  The exchange lists of
 * hk_set_node_list(0 / 1 / 2) must then be STATIC and cover every node that can ever join a surface — all nodes of the
 * instances in contact — because nothing rebuilds them (nodes without a contact slot export zeros and ignore imports).
 * Every step the engine logs its own deletions as global ids; with a communicator (hk_comm_init) it all-gathers
 * {count, ids[max_deleted_per_step]} of every rank on its stream and one thread replays them in ascending global order
 * (ranks own contiguous ascending element blocks) — pair lists, contact slots and special-node table grow on the
 * device, and hk_step_enqueue(t, n > 1) needs no host in between.  Without a communicator the host still gathers the
 * ids and calls hk_apply_deleted, which then replays them on the device.  A step that deletes more than
 * max_deleted_per_step elements on one rank is reported by hk_sync as an error (HK_ERR_STATE). */
int hk_comm_erosion(hk_engine* e, int32_t max_deleted_per_step);
/* Without hk_set_global_maps (single-domain engine) hk_apply_deleted is the RESTART hook: ids are the engine's own
 * 1-based element ids in their original deletion order (hk_deleted_ids of the checkpointed run); the exposed-face
 * updates are replayed and the ids are recorded as already deleted.  State itself comes back through
 * hk_upload_state (hakai_fem_b200/checkpoint.py). */

#ifdef __cplusplus
}
#endif
#endif /* HAKAI_B200_H */
