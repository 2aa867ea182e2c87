// hk_engine.cu — host side of libhakai_b200.so: the C ABI of include/hakai_b200.h.
//
// Owns all device memory, converts the reference's Julia layouts (1-based Int64, AoS (6,nip)) to the
// device layouts of hk_common.h once at hk_finalize / hk_upload_state and back at hk_download, and
// drives the per-step kernel sequence of J2:487-951:
//     [contact kernels] -> nodal kernel (gather Q, update, BC, kinematics) -> element kernel
// The exposed-face update after element deletion (add_surface_triangle, J2:767-804, 2167-2245) runs
// here on the host with a sorted face table instead of the reference's O(6F) scan per deleted element.
#include <algorithm>
#include <array>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/hakai_b200.h"
#include "hk_common.h"

#ifdef HK_EMU
#define HKAPI(name) hke_##name
#else
#define HKAPI(name) hk_##name
#endif

namespace {

struct MaterialH {
    double young, poisson, density;
    std::vector<double> plastic, Hd, ductile;   // column-major as given
    int64_t npp, nd;
};
struct BCH {
    std::vector<std::vector<int64_t>> dof;
    std::vector<double> value;
    std::vector<double> a_t, a_v;
};
struct ICH {
    std::vector<std::vector<int64_t>> dof;
    std::vector<double> value;
};
struct InstanceH {
    int64_t node_offset, nNode, element_offset, nElement;
    std::vector<int64_t> surfaces, eleid;        // (F,4) column-major part-local 1-based; (F)
    std::vector<int> twin;                       // (F) from hk_build_contact (else built at the first step)
    bool table_built = false;
    std::vector<std::array<int64_t, 4>> key;     // sorted 4-tuples
    std::vector<int64_t> order;                  // face ids sorted by (key, id)
};
struct PairH {
    int64_t i_instance, j_instance;
    std::vector<int> nodes_i, nodes_j, t0, t1, t2, tele;   // 0-based; mirror of the device lists (see contact_refresh_host)
    std::unordered_set<int> set_i, set_j;
    double young;
    HkPairDev dev;
    bool dev_valid = false;
};
struct InstDevH {                   // device tables of one instance for erode_element
    int* surf = nullptr;
    int* feleid = nullptr;
    int* twin = nullptr;
};
struct HaloNbr {
    std::vector<int> nodes;        // local 0-based node ids
    std::vector<int> slots;        // index into the dense halo node list
    int* d_nodes = nullptr;
    int* d_slots = nullptr;
    double* send = nullptr;        // caller-owned device buffers
    double* recv = nullptr;
    int64_t rank = -1;             // global rank of the neighbour (hk_set_halo_ranks); -1: unknown
    bool own_buffers = false;      // send/recv allocated by the engine (hk_comm_init) rather than bound by the host
};
struct TimedEvent {
#ifndef HK_EMU
    cudaEvent_t a, b;
#endif
    int kind;
};

}  // namespace

struct hk_engine {
    hk_params prm;
    std::string err;
    bool finalized = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t nNode = 0, nElement = 0;
    std::vector<double> coordmat, mass;
    std::vector<int> conn;                       // [e*8+a] 0-based
    std::vector<int> emat;                       // 0-based
    std::vector<int64_t> einst;                  // 1-based
    std::vector<MaterialH> materials;
    std::vector<BCH> bcs;
    std::vector<ICH> ics;
    std::vector<InstanceH> instances;
    std::vector<PairH> pairs;
    std::vector<HaloNbr> halo;
    int n_halo_nodes = 0;
    int* d_halo_list = nullptr;    // node id of every halo slot (nodal kernel mode 2)
    double* d_halo_own = nullptr;  // [n_halo*3] this rank's own partial force of every halo slot
    int64_t my_rank = -1;          // global rank of this engine (hk_set_halo_ranks); -1: unknown
    // NCCL inside the library (hk_comm_init): communicator, side stream for the exchange, ordering events
    void* comm = nullptr;
    int comm_world = 0;
    // contact exchange inside the library (hk_comm_contact): padded export block, gathered blocks of all ranks, limbs
    int64_t cx_maxlen = 0;
    double* cx_send = nullptr;
    double* cx_all = nullptr;
    long long* cx_limbs = nullptr;
#ifndef HK_EMU
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_pack = nullptr, ev_comm = nullptr;
#endif
    // multi-GPU node lists: contact 0 own-export, 1 ghost-import, 2 surface nodes (force exchange);
    // ghost-element mode 3 state-export, 4 state-import
    std::vector<int> node_list[5];
    int* d_node_list[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    long long* d_import_src = nullptr;
    // exposed-face update on the device (single-domain engines whose contact surfaces can erode)
    bool dev_erosion = false;      // tables built, erode_element runs after every step
    bool erosion_checked = false;  // decided at the first step (multi-GPU drivers call hk_set_global_maps after finalize)
    bool contact_host_stale = false;   // device lists may have grown since the host mirrors were read
    std::vector<char> inst_erodes; // per instance: has surfaces and an element whose material can fail
    std::vector<InstDevH> inst_dev;
    HkErodeDev er;
    // v0.0.1 penetration clamp: d_node ping-pong [2][slot capacity], d_max ping-pong [2] (bit patterns), current index
    double* d_dnode = nullptr;
    unsigned long long* d_dmax = nullptr;
    size_t dnode_cap = 0;
    int clamp_cur = 0;
    bool contact_done = false;     // hk_contact_enqueue already ran the contact pass of the next step
    bool frame_next = false;       // hk_mark_frame: the next asynchronous step stores integ_triax_stress
    // multi-GPU erosion: instance tables are GLOBAL; these map global (1-based) ids to engine-local 0-based ids or -1
    std::vector<int> g_node_map, g_elem_map;
    // step replay by CUDA graph (single-domain engines on their own stream): the launches of ONE step captured once and
    // replayed for every further step; the step number lives on the device (d.t_dev).  Small decks are bound by the
    // host's launch rate (the reference's example decks: 16 launches and 190 us of host time per step of 25 us of kernels).
    long long* d_step = nullptr;
#ifndef HK_EMU
    cudaGraphExec_t step_graph = nullptr;
#endif
    bool capturing = false;
    bool graph_dirty = true;       // set by every device (re)allocation and by anything else a captured launch depends on
    bool graph_off = false;        // HK_STEP_GRAPH=0, or a capture failed: plain launches from then on
    long long graph_launches = 0;  // kernel launches inside one replay
    // hk_comm_erosion: deletions of all ranks replayed on the device (static exchange lists over every candidate node)
    bool xerode = false;
    int xe_cap = 0;                // deletions per rank and step the all-gather carries
    long long* xe_send = nullptr;  // [xe_cap + 1]: n, global 0-based ids of this rank's deletions of the step
    long long* xe_all = nullptr;   // [world][xe_cap + 1]
    int* d_e_l2g = nullptr;        // engine element -> global 0-based id
    int* d_node_key = nullptr;     // engine node -> global 0-based id
    std::vector<int64_t> g_einst;  // instance of every GLOBAL element
    // hk_node_output work buffers, allocated at the first call and kept (cudaMalloc/cudaFree of ~4 GB per frame costs
    // more than the averaging itself)
    double* no_emean = nullptr;    // [14][nEp]
    double* no_out = nullptr;      // [16][nNode]
    unsigned long long* d_summary = nullptr;   // hk_state_summary scratch
    int64_t begun_t = -1;          // step opened by hk_step_begin
    // special nodes (host mirror)
    std::vector<int> spec_idx_h;
    std::vector<HkSpecialNode> spec_h;
    int n_contact_slots = 0;
    size_t spec_cap = 0, cacc_cap = 0;
    HkDev d;
    HkContactParams cp;
    bool any_ductile = false;
    bool velo_current = true;      // d.velo holds the current velocity
    bool triax_current = true;     // d.triax matches the state
    int use_Q0 = 0;
    double dt2 = 0, dt2p = 0;
    std::vector<void*> allocs;
    double* staging = nullptr;     // device staging for layout transposes
    size_t staging_doubles = 0;
    // double-buffered transfer of the Gauss-point state (hk_upload_state / hk_download): two staging blocks, a copy
    // stream beside the engine's stream, events instead of host synchronisation per chunk
    double* stage2[2] = {nullptr, nullptr};
    size_t stage2_doubles = 0;
#ifndef HK_EMU
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
#endif
    std::vector<int64_t> deleted_all;
    std::vector<int64_t> deleted_step;          // step t of every entry of deleted_all (0: replayed by hk_apply_deleted)
    size_t deleted_reported = 0;
    int del_seen = 0;
    int64_t n_launch = 0, n_steps = 0;
    bool profiling = false;
    std::vector<TimedEvent> events;
    double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // kinds 0-3 as hk_profile_read; 4 halo exchange; 5 deletion pass
    int64_t prof_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static std::string g_create_err;

// Engines on different devices may live in one process: every ABI entry that touches the device selects the
// engine's device first (kernel attributes are set per launch, constant tables uploaded per engine at hk_finalize).
#ifndef HK_EMU
#define HK_DEVICE(e) do { cudaError_t rd_ = cudaSetDevice((e)->prm.device); if (rd_ != cudaSuccess) return cuda_fail(e, (int)rd_, "cudaSetDevice"); } while (0)
#else
#define HK_DEVICE(e) do { } while (0)
#endif

static int fail(hk_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg; else g_create_err = msg;
    return code;
}
static int cuda_fail(hk_engine* e, int rc, const char* what) {
    return fail(e, HK_ERR_CUDA, std::string(what) + ": " + hkp::error_string(rc));
}
#define CK(call) do { int rc_ = (int)(call); if (rc_) return cuda_fail(e, rc_, #call); } while (0)

template <class T>
static int dalloc(hk_engine* e, T** p, size_t count) {
    void* q = nullptr;
    int rc = hkp::dev_malloc(&q, count * sizeof(T));
    if (rc) return cuda_fail(e, rc, "device allocation");
    e->allocs.push_back(q);
    e->graph_dirty = true;
    *p = (T*)q;
    return 0;
}
static void dfree(hk_engine* e, void* p) {
    e->graph_dirty = true;
    if (!p) return;
    auto it = std::find(e->allocs.begin(), e->allocs.end(), p);
    if (it != e->allocs.end()) e->allocs.erase(it);
    hkp::dev_free(p);
}
template <class T>
static int upload(hk_engine* e, T* dst, const std::vector<T>& src) {
    if (src.empty()) return 0;
    CK(hkp::h2d(dst, src.data(), src.size() * sizeof(T), e->stream));
    return 0;
}

// --------------------------------------------------------------------------------- profiling helpers
static void prof_begin(hk_engine* e, int kind, cudaStream_t on = nullptr, bool other_stream = false) {
#ifndef HK_EMU
    if (!e->profiling) return;
    TimedEvent t;
    t.kind = kind;
    cudaEventCreate(&t.a);
    cudaEventCreate(&t.b);
    cudaEventRecord(t.a, other_stream ? on : e->stream);
    e->events.push_back(t);
#else
    (void)e; (void)kind; (void)on; (void)other_stream;
#endif
}
static void prof_end(hk_engine* e, cudaStream_t on = nullptr, bool other_stream = false) {
#ifndef HK_EMU
    if (!e->profiling) return;
    cudaEventRecord(e->events.back().b, other_stream ? on : e->stream);
#else
    (void)e; (void)on; (void)other_stream;
#endif
}
static void prof_collect(hk_engine* e) {
#ifndef HK_EMU
    if (e->events.empty()) return;
    cudaStreamSynchronize(e->stream);
    if (e->comm_stream) cudaStreamSynchronize(e->comm_stream);
    for (size_t i = 0; i < e->events.size(); ++i) {
        TimedEvent& t = e->events[i];
        float ms = 0;
        cudaEventElapsedTime(&ms, t.a, t.b);
        e->prof_ms[t.kind] += ms;
        e->prof_n[t.kind] += 1;
        if (i + 1 < e->events.size() && t.kind != 4 && e->events[i + 1].kind != 4) {   // kind 6: idle time between the
            float gap = 0;                                                                // profiled launches of the main stream
            if (cudaEventElapsedTime(&gap, t.b, e->events[i + 1].a) == cudaSuccess && gap > 0) { e->prof_ms[6] += gap; e->prof_n[6] += 1; }
        }
    }
    for (auto& t : e->events) {
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    e->events.clear();
#else
    (void)e;
#endif
}

// --------------------------------------------------------------------------------- Pusai (J2:1895-1943)
static void cal_Pusai_hexa(double P[8][3][8]) {
    static const double dm[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                    {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
    const double g = 1.0 / std::sqrt(3.0);
    const double gc[8][3] = {{-g, -g, -g}, {-g, -g, g}, {-g, g, -g}, {-g, g, g},
                             {g, -g, -g},  {g, -g, g},  {g, g, -g},  {g, g, g}};
    for (int k = 0; k < 8; ++k)
        for (int i = 0; i < 8; ++i) {
            P[k][0][i] = 1.0 / 8.0 * dm[i][0] * (1.0 + gc[k][1] * dm[i][1]) * (1.0 + gc[k][2] * dm[i][2]);
            P[k][1][i] = 1.0 / 8.0 * dm[i][1] * (1.0 + gc[k][0] * dm[i][0]) * (1.0 + gc[k][2] * dm[i][2]);
            P[k][2][i] = 1.0 / 8.0 * dm[i][2] * (1.0 + gc[k][0] * dm[i][0]) * (1.0 + gc[k][1] * dm[i][1]);
        }
}

// --------------------------------------------------------------------------------- contact pair device arrays
static bool inst_can_erode(const hk_engine* e, int64_t inst) {
    return e->dev_erosion && inst >= 1 && inst <= (int64_t)e->inst_erodes.size() && e->inst_erodes[inst - 1];
}

// (re)creates the device arrays of a pair from the host mirrors.  With device-side erosion the lists are allocated at
// their worst-case length (every node of an eroding instance exposed, two triangles per face of its elements), so
// erode_element can append without the host.
static int pair_upload(hk_engine* e, PairH& p) {
    HkPairDev& D = p.dev;
    if (p.dev_valid) {
        dfree(e, D.nodes_i); dfree(e, D.nodes_j); dfree(e, D.t0); dfree(e, D.t1); dfree(e, D.t2); dfree(e, D.tele);
        dfree(e, D.bbox); dfree(e, D.cell_i); dfree(e, D.head); dfree(e, D.next); dfree(e, D.dyn); dfree(e, D.cand);
        dfree(e, D.in_i); dfree(e, D.in_j);
        p.dev_valid = false;
    }
    std::memset(&D, 0, sizeof(D));
    const int nn_i = (int)p.nodes_i.size(), nn_j = (int)p.nodes_j.size(), nTri = (int)p.t0.size();
    D.i_instance = (int)p.i_instance; D.j_instance = (int)p.j_instance;
    D.self = p.i_instance == p.j_instance;
    D.young = p.young;
    const bool ei = inst_can_erode(e, p.i_instance), ej = inst_can_erode(e, p.j_instance);
    long long cap_i = nn_i, cap_j = nn_j, cap_tri = nTri;
    if (ei) cap_i += e->instances[p.i_instance - 1].nNode;
    if (ej && !D.self) { cap_j += e->instances[p.j_instance - 1].nNode; cap_tri += 12 * e->instances[p.j_instance - 1].nElement; }
    if (cap_i >= (1ll << 30) || cap_j >= (1ll << 30) || cap_tri >= (1ll << 31) - 1) return fail(e, HK_ERR_UNSUPPORTED, "contact pair too large");
    D.cap_i = (int)cap_i; D.cap_j = (int)cap_j; D.cap_tri = (int)cap_tri;
    auto pow2 = [](long long n) { int nb = 64; while (nb < 2 * n) nb <<= 1; return nb; };
    D.cap_bucket = pow2(cap_i);
    int rc;
    if ((rc = dalloc(e, &D.nodes_i, (size_t)D.cap_i))) return rc;
    if ((rc = dalloc(e, &D.nodes_j, (size_t)D.cap_j))) return rc;
    if ((rc = dalloc(e, &D.t0, (size_t)D.cap_tri))) return rc;
    if ((rc = dalloc(e, &D.t1, (size_t)D.cap_tri))) return rc;
    if ((rc = dalloc(e, &D.t2, (size_t)D.cap_tri))) return rc;
    if ((rc = dalloc(e, &D.tele, (size_t)D.cap_tri))) return rc;
    if ((rc = dalloc(e, &D.bbox, (size_t)12))) return rc;
    if ((rc = dalloc(e, &D.cell_i, (size_t)3 * D.cap_i))) return rc;
    if ((rc = dalloc(e, &D.head, (size_t)D.cap_bucket))) return rc;
    if ((rc = dalloc(e, &D.next, (size_t)D.cap_i))) return rc;
    if ((rc = dalloc(e, &D.cand, (size_t)D.cap_tri))) return rc;
    if ((rc = dalloc(e, &D.dyn, (size_t)1))) return rc;
    if ((rc = upload(e, D.nodes_i, p.nodes_i))) return rc;
    if ((rc = upload(e, D.nodes_j, p.nodes_j))) return rc;
    if ((rc = upload(e, D.t0, p.t0))) return rc;
    if ((rc = upload(e, D.t1, p.t1))) return rc;
    if ((rc = upload(e, D.t2, p.t2))) return rc;
    if ((rc = upload(e, D.tele, p.tele))) return rc;
    const HkPairDyn dyn = {nn_i, nn_j, nTri, pow2(nn_i), 0};
    CK(hkp::h2d(D.dyn, &dyn, sizeof(dyn), e->stream));
    if (ei || ej) {                                   // membership bytes for the device-side `unique!` (J2:786, 791)
        std::vector<unsigned char> in(e->nNode, 0);
        for (int n : p.nodes_i) in[n] = 1;
        if ((rc = dalloc(e, &D.in_i, (size_t)e->nNode))) return rc;
        if ((rc = upload(e, D.in_i, in))) return rc;
        std::fill(in.begin(), in.end(), 0);
        for (int n : p.nodes_j) in[n] = 1;
        if ((rc = dalloc(e, &D.in_j, (size_t)e->nNode))) return rc;
        if ((rc = upload(e, D.in_j, in))) return rc;
    }
    p.dev_valid = true;
    return 0;
}

// special-node bookkeeping ---------------------------------------------------------------------
static int spec_of(hk_engine* e, int node) {
    int s = e->spec_idx_h[node];
    if (s < 0) {
        HkSpecialNode sn;
        sn.bc_entry[0] = sn.bc_entry[1] = sn.bc_entry[2] = -1;
        sn.contact_slot = -1;
        sn.halo_slot = -1;
        sn.pad = 0;
        s = (int)e->spec_h.size();
        e->spec_h.push_back(sn);
        e->spec_idx_h[node] = s;
    }
    return s;
}
static void ensure_contact_slot(hk_engine* e, int node, std::vector<int>* touched) {
    int s = spec_of(e, node);
    if (e->spec_h[s].contact_slot < 0) {
        e->spec_h[s].contact_slot = e->n_contact_slots++;
        if (touched) touched->push_back(node);
    }
}
static int spec_upload(hk_engine* e, const std::vector<int>* touched_nodes) {
    // (re)upload the special-node table; grow device arrays when needed
    if (e->spec_h.size() > e->spec_cap) {
        dfree(e, e->d.spec);
        e->spec_cap = std::max<size_t>(64, e->spec_h.size() * 2);
        int rc = dalloc(e, &e->d.spec, e->spec_cap);
        if (rc) return rc;
    }
    int rc = upload(e, e->d.spec, e->spec_h);
    if (rc) return rc;
    if ((size_t)e->n_contact_slots > e->cacc_cap) {
        dfree(e, e->d.cacc);
        e->cacc_cap = std::max<size_t>(64, (size_t)e->n_contact_slots * 2);
        rc = dalloc(e, &e->d.cacc, e->cacc_cap * 6);
        if (rc) return rc;
        CK(hkp::dev_memset(e->d.cacc, 0, e->cacc_cap * 6 * sizeof(unsigned long long), e->stream));
    }
    if (touched_nodes) {
        for (int n : *touched_nodes) CK(hkp::h2d(e->d.spec_idx + n, &e->spec_idx_h[n], sizeof(int), e->stream));
    } else {
        rc = upload(e, e->d.spec_idx, e->spec_idx_h);
        if (rc) return rc;
    }
    if (e->dev_erosion) {                     // lengths live on the device in this mode
        const int ns_h = (int)e->spec_h.size(), nl_h = e->n_contact_slots;
        CK(hkp::h2d(e->er.n_spec, &ns_h, sizeof(int), e->stream));
        CK(hkp::h2d(e->er.n_slots, &nl_h, sizeof(int), e->stream));
    }
    return 0;
}

static int erosion_upload_descriptors(hk_engine* e);
int hk_setup_instance_faces(const int* d_conn, long long nEp, const double* d_X, long long node_offset, long long element_offset,
                            long long nElement, cudaStream_t s, std::vector<int>& surf, std::vector<int>& twin,
                            std::vector<int>& exterior_ids);

// --------------------------------------------------------------------------------- exposed faces (A10)
static void build_face_table(InstanceH& I) {
    const int64_t F = 6 * I.nElement;
    I.key.resize(F);
    for (int64_t j = 0; j < F; ++j) {
        std::array<int64_t, 4> k = {I.surfaces[j], I.surfaces[j + F], I.surfaces[j + 2 * F], I.surfaces[j + 3 * F]};
        std::sort(k.begin(), k.end());
        I.key[j] = k;
    }
    I.order.resize(F);
    for (int64_t j = 0; j < F; ++j) I.order[j] = j;
    std::sort(I.order.begin(), I.order.end(), [&](int64_t a, int64_t b) {
        if (I.key[a] != I.key[b]) return I.key[a] < I.key[b];
        return a < b;
    });
    I.table_built = true;
}

// add_surface_triangle (J2:2167-2245): for each face of the deleted element, the first face (in face-id
// order) of ANOTHER element with the same node set becomes exposed.
static void add_surface_triangle(InstanceH& I, int64_t ele_id, std::vector<int64_t>& tri, std::vector<int64_t>& tri_ele,
                                 std::vector<int64_t>& nodes) {
    if (!I.table_built) build_face_table(I);
    const int64_t F = 6 * I.nElement;
    for (int j = 0; j < 6; ++j) {
        const int64_t fj = 6 * (ele_id - 1) + j;
        const std::array<int64_t, 4>& kj = I.key[fj];
        auto lo = std::lower_bound(I.order.begin(), I.order.end(), fj, [&](int64_t a, int64_t) { return I.key[a] < kj; });
        for (auto it = lo; it != I.order.end() && I.key[*it] == kj; ++it) {
            const int64_t k = *it;
            if (I.eleid[k] == ele_id) continue;
            const int64_t s[4] = {I.surfaces[k], I.surfaces[k + F], I.surfaces[k + 2 * F], I.surfaces[k + 3 * F]};
            tri.push_back(s[0]); tri.push_back(s[1]); tri.push_back(s[2]);
            tri.push_back(s[2]); tri.push_back(s[3]); tri.push_back(s[0]);
            tri_ele.push_back(I.eleid[k]);
            tri_ele.push_back(I.eleid[k]);
            break;
        }
    }
    nodes = tri;
    std::sort(nodes.begin(), nodes.end());
    nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
}

// `deleted`: 1-based element ids — engine-local ids normally; GLOBAL ids when the multi-GPU maps are set
// (hk_set_global_maps): then the instance tables are global, nodes are translated through g_node_map and only
// triangles of locally owned elements are kept (another rank owns the others).
static int update_surfaces(hk_engine* e, const std::vector<int64_t>& deleted) {
    const bool global = !e->g_node_map.empty();
    std::vector<char> changed(e->pairs.size(), 0);
    std::vector<int> touched;
    for (int64_t gid : deleted) {
        const int64_t instance_id = global ? e->g_einst[gid - 1] : e->einst[gid - 1];
        if (instance_id < 1 || instance_id > (int64_t)e->instances.size()) continue;
        InstanceH& I = e->instances[instance_id - 1];
        if (I.surfaces.empty()) continue;
        std::vector<int64_t> tri, tri_ele, nodes;
        add_surface_triangle(I, gid - I.element_offset, tri, tri_ele, nodes);
        auto node_of = [&](int64_t part_local) -> int {
            const int64_t g = part_local + I.node_offset;           // 1-based (global) node id
            return global ? e->g_node_map[g - 1] : (int)(g - 1);
        };
        for (size_t c = 0; c < e->pairs.size(); ++c) {
            PairH& p = e->pairs[c];
            if (p.i_instance == instance_id) {                  // J2:784-787
                for (int64_t nl : nodes) {
                    const int g = node_of(nl);
                    if (g < 0) return fail(e, HK_ERR_STATE, "exposed node is not present on this rank (ghost set too small)");
                    if (p.set_i.insert(g).second) { p.nodes_i.push_back(g); ensure_contact_slot(e, g, &touched); changed[c] = 1; }
                }
            } else if (p.j_instance == instance_id) {           // J2:789-797
                for (int64_t nl : nodes) {
                    const int g = node_of(nl);
                    if (g < 0) return fail(e, HK_ERR_STATE, "exposed node is not present on this rank (ghost set too small)");
                    if (p.set_j.insert(g).second) { p.nodes_j.push_back(g); ensure_contact_slot(e, g, &touched); changed[c] = 1; }
                }
                for (size_t r = 0; r < tri_ele.size(); ++r) {
                    const int64_t ge = tri_ele[r] + I.element_offset;   // 1-based (global) element id
                    const int le = global ? e->g_elem_map[ge - 1] : (int)(ge - 1);
                    if (le < 0) continue;                               // owned by another rank
                    p.t0.push_back(node_of(tri[3 * r + 0]));
                    p.t1.push_back(node_of(tri[3 * r + 1]));
                    p.t2.push_back(node_of(tri[3 * r + 2]));
                    p.tele.push_back(le);
                    changed[c] = 1;
                }
            }
        }
    }
    bool any_changed = false;
    for (size_t c = 0; c < e->pairs.size(); ++c)
        if (changed[c]) { int rc = pair_upload(e, e->pairs[c]); if (rc) return rc; any_changed = true; }
    if (!touched.empty()) { int rc = spec_upload(e, &touched); if (rc) return rc; }
    if (any_changed && e->dev_erosion) { int rc = erosion_upload_descriptors(e); if (rc) return rc; }
    return 0;
}

// ---- exposed faces on the device --------------------------------------------------------------------------------
// twin[f]: the face add_surface_triangle (J2:2167-2245) would find for face f of a deleted element — the first face, in
// face-id order, of ANOTHER element with the same node set — or -1.  Sort-based: faces are ordered by (hash of the
// sorted node 4-tuple, face id); equal tuples are adjacent, hash collisions are resolved by comparing the tuples.
static void build_twin_table(const InstanceH& I, std::vector<int>& twin) {
    const int64_t F = 6 * I.nElement;
    struct Rec { uint64_t key; int64_t face; };
    std::vector<Rec> rec(F);
    std::vector<std::array<int64_t, 4>> tup(F);
    for (int64_t j = 0; j < F; ++j) {
        std::array<int64_t, 4> k = {I.surfaces[j], I.surfaces[j + F], I.surfaces[j + 2 * F], I.surfaces[j + 3 * F]};
        std::sort(k.begin(), k.end());
        tup[j] = k;
        uint64_t h = 0x9e3779b97f4a7c15ull;
        for (int a = 0; a < 4; ++a) { h ^= (uint64_t)k[a] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); h *= 0xff51afd7ed558ccdull; }
        rec[j] = {h, j};
    }
    std::sort(rec.begin(), rec.end(), [](const Rec& a, const Rec& b) { return a.key != b.key ? a.key < b.key : a.face < b.face; });
    twin.assign(F, -1);
    for (int64_t lo = 0; lo < F;) {
        int64_t hi = lo + 1;
        while (hi < F && rec[hi].key == rec[lo].key) ++hi;
        if (hi - lo > 1)
            for (int64_t a = lo; a < hi; ++a) {
                const int64_t f = rec[a].face;
                for (int64_t b = lo; b < hi; ++b) {           // ascending face id
                    const int64_t k = rec[b].face;
                    if (I.eleid[k] == I.eleid[f] || tup[k] != tup[f]) continue;
                    twin[f] = (int)k;
                    break;
                }
            }
        lo = hi;
    }
}

static int erosion_upload_descriptors(hk_engine* e) {
    std::vector<HkPairDev> pd;
    for (const PairH& p : e->pairs) pd.push_back(p.dev);
    return upload(e, e->er.pairs, pd);
}

// Decided at the first step: a single-domain engine (no hk_set_global_maps) with contact and a material that can fail
// keeps its contact surfaces current on the device.  Builds the per-instance twin tables, re-creates the pair lists
// at worst-case capacity and moves the special-node / accumulator bookkeeping to device counters.
static int ensure_erosion(hk_engine* e) {
    if (e->erosion_checked) return 0;
    e->erosion_checked = true;
    const bool contact_on = e->prm.contact_flag >= 1 && !e->pairs.empty();
    const bool global = !e->g_node_map.empty();       // partitioned mesh: GLOBAL instance tables (hk_set_global_maps)
    if (!contact_on || !e->any_ductile) return 0;
    if (global && !e->xerode) return 0;               // the host driver replays deletions through hk_apply_deleted
    const size_t nI = e->instances.size();
    if (nI == 0 || nI >= 65535) return global ? fail(e, HK_ERR_STATE, "hk_comm_erosion: no instance tables") : 0;
    std::vector<char> used(nI, 0);
    for (const PairH& p : e->pairs) {
        if (p.i_instance >= 1 && p.i_instance <= (int64_t)nI) used[p.i_instance - 1] = 1;
        if (p.j_instance >= 1 && p.j_instance <= (int64_t)nI) used[p.j_instance - 1] = 1;
    }
    e->inst_erodes.assign(nI, 0);
    bool any = false;
    if (global) {                 // a rank sees the materials of its own elements only: any instance in contact may erode
        for (size_t i = 0; i < nI; ++i)
            if (used[i] && !e->instances[i].surfaces.empty()) { e->inst_erodes[i] = 1; any = true; }
    } else {
        for (int64_t el = 0; el < e->nElement; ++el) {
            const int64_t inst = e->einst[el];
            if (inst < 1 || inst > (int64_t)nI || !used[inst - 1] || e->instances[inst - 1].surfaces.empty()) continue;
            if (e->materials[e->emat[el]].nd > 0) { e->inst_erodes[inst - 1] = 1; any = true; }
        }
    }
    if (!any) { e->xerode = false; return 0; }
    const int64_t nE_all = global ? (int64_t)e->g_elem_map.size() : e->nElement;
    const int64_t nN_all = global ? (int64_t)e->g_node_map.size() : e->nNode;
    int rc;
    HkErodeDev& E = e->er;
    std::memset(&E, 0, sizeof(E));
    E.n_inst = (int)nI;
    E.n_pair = (int)e->pairs.size();
    e->inst_dev.assign(nI, InstDevH());
    std::vector<HkInstDev> idv(nI);
    long long extra_nodes = 0;
    for (size_t i = 0; i < nI; ++i) {
        const InstanceH& I = e->instances[i];
        HkInstDev& D = idv[i];
        std::memset(&D, 0, sizeof(D));
        D.element_offset = I.element_offset;
        D.nElement = I.nElement;
        if (!e->inst_erodes[i]) continue;
        if (I.element_offset < 0 || I.element_offset + I.nElement > nE_all || I.node_offset < 0 ||
            I.node_offset + I.nNode > nN_all)
            return fail(e, HK_ERR_ARG, "instance range outside the mesh");
        const int64_t F = 6 * I.nElement;
        std::vector<int> twin, surf((size_t)4 * F), fele(F);
        if ((int64_t)I.twin.size() == F) twin = I.twin; else build_twin_table(I, twin);
        for (int64_t j = 0; j < 4 * F; ++j) {
            const int64_t v = I.surfaces[j];
            if (v < 1 || v > I.nNode) return fail(e, HK_ERR_ARG, "instance face node out of range");
            const int64_t g = v + I.node_offset - 1;
            surf[j] = global ? e->g_node_map[g] : (int)g;
            if (surf[j] < 0) return fail(e, HK_ERR_STATE, "node of an instance in contact is not present on this rank (ghost set too small)");
        }
        for (int64_t j = 0; j < F; ++j) {
            const int64_t v = I.eleid[j];
            if (v < 1 || v > I.nElement) return fail(e, HK_ERR_ARG, "instance face element out of range");
            const int64_t g = v + I.element_offset - 1;
            fele[j] = global ? e->g_elem_map[g] : (int)g;       // -1: another rank owns the element
        }
        InstDevH& H = e->inst_dev[i];
        if ((rc = dalloc(e, &H.surf, surf.size()))) return rc;
        if ((rc = dalloc(e, &H.feleid, fele.size()))) return rc;
        if ((rc = dalloc(e, &H.twin, twin.size()))) return rc;
        if ((rc = upload(e, H.surf, surf))) return rc;
        if ((rc = upload(e, H.feleid, fele))) return rc;
        if ((rc = upload(e, H.twin, twin))) return rc;
        D.surf = H.surf; D.feleid = H.feleid; D.twin = H.twin;
        extra_nodes += I.nNode;
    }
    if ((rc = dalloc(e, &E.inst, nI))) return rc;
    if ((rc = upload(e, E.inst, idv))) return rc;
    {
        std::vector<unsigned short> ei(nE_all);
        for (int64_t el = 0; el < nE_all; ++el) {
            const int64_t v = global ? e->g_einst[el] : e->einst[el];
            ei[el] = (unsigned short)((v >= 1 && v <= (int64_t)nI) ? v : 0);
        }
        if ((rc = dalloc(e, &E.einst, ei.size()))) return rc;
        if ((rc = upload(e, E.einst, ei))) return rc;
    }
    if (global) {
        std::vector<int> l2g(e->nElement, 0), key(e->nNode, 0);
        for (int64_t g = 0; g < nE_all; ++g) if (e->g_elem_map[g] >= 0) l2g[e->g_elem_map[g]] = (int)g;
        for (int64_t g = 0; g < nN_all; ++g) if (e->g_node_map[g] >= 0) key[e->g_node_map[g]] = (int)g;
        if ((rc = dalloc(e, &e->d_e_l2g, l2g.size()))) return rc;
        if ((rc = upload(e, e->d_e_l2g, l2g))) return rc;
        if ((rc = dalloc(e, &e->d_node_key, key.size()))) return rc;
        if ((rc = upload(e, e->d_node_key, key))) return rc;
        E.node_key = e->d_node_key;
        const int world = e->comm ? e->comm_world : 1;
        if ((rc = dalloc(e, &e->xe_send, (size_t)e->xe_cap + 1))) return rc;
        if ((rc = dalloc(e, &e->xe_all, (size_t)world * (e->xe_cap + 1)))) return rc;
        CK(hkp::dev_memset(e->xe_send, 0, ((size_t)e->xe_cap + 1) * sizeof(long long), e->stream));
    }
    if ((rc = dalloc(e, &E.n_spec, (size_t)1))) return rc;
    if ((rc = dalloc(e, &E.n_slots, (size_t)1))) return rc;
    if ((rc = dalloc(e, &E.overflow, (size_t)1))) return rc;
    if ((rc = dalloc(e, &E.pairs, e->pairs.size()))) return rc;
    CK(hkp::dev_memset(E.overflow, 0, sizeof(int), e->stream));
    // special-node table and accumulators at worst-case capacity, lengths on the device from now on
    e->dev_erosion = true;
    E.spec_cap = (int)std::min<long long>((long long)e->spec_h.size() + extra_nodes, e->nNode);
    E.slot_cap = (int)std::min<long long>((long long)e->n_contact_slots + extra_nodes, e->nNode);
    {
        HkSpecialNode* ns = nullptr;
        unsigned long long* nc = nullptr;
        if ((rc = dalloc(e, &ns, (size_t)E.spec_cap))) return rc;
        if ((rc = dalloc(e, &nc, (size_t)E.slot_cap * 6))) return rc;
        dfree(e, e->d.spec);
        dfree(e, e->d.cacc);
        e->d.spec = ns; e->d.cacc = nc;
        e->spec_cap = (size_t)E.spec_cap;
        e->cacc_cap = (size_t)E.slot_cap;
        if ((rc = upload(e, e->d.spec, e->spec_h))) return rc;
        CK(hkp::dev_memset(e->d.cacc, 0, (size_t)E.slot_cap * 6 * sizeof(unsigned long long), e->stream));
        const int ns_h = (int)e->spec_h.size(), nl_h = e->n_contact_slots;
        CK(hkp::h2d(E.n_spec, &ns_h, sizeof(int), e->stream));
        CK(hkp::h2d(E.n_slots, &nl_h, sizeof(int), e->stream));
    }
    for (PairH& p : e->pairs) if ((rc = pair_upload(e, p))) return rc;
    return erosion_upload_descriptors(e);
}

// Host mirrors of everything erode_element may have grown: pair lists, special-node table, slot count.
static int contact_refresh_host(hk_engine* e) {
    if (!e->dev_erosion || !e->contact_host_stale) return 0;
    CK(hkp::sync(e->stream));
    bool grew = false;
    for (PairH& p : e->pairs) {
        HkPairDyn dyn;
        CK(hkp::d2h(&dyn, p.dev.dyn, sizeof(dyn), e->stream));
        auto tail = [&](std::vector<int>& v, const int* dev, int n) -> int {
            const size_t old = v.size();
            if ((size_t)n <= old) return 0;
            v.resize(n);
            grew = true;
            return hkp::d2h(v.data() + old, dev + old, ((size_t)n - old) * sizeof(int), e->stream);
        };
        const size_t oi = p.nodes_i.size(), oj = p.nodes_j.size();
        CK(tail(p.nodes_i, p.dev.nodes_i, dyn.nn_i));
        CK(tail(p.nodes_j, p.dev.nodes_j, dyn.nn_j));
        CK(tail(p.t0, p.dev.t0, dyn.nTri));
        CK(tail(p.t1, p.dev.t1, dyn.nTri));
        CK(tail(p.t2, p.dev.t2, dyn.nTri));
        CK(tail(p.tele, p.dev.tele, dyn.nTri));
        for (size_t k = oi; k < p.nodes_i.size(); ++k) p.set_i.insert(p.nodes_i[k]);
        for (size_t k = oj; k < p.nodes_j.size(); ++k) p.set_j.insert(p.nodes_j[k]);
    }
    if (grew) {
        int ns = 0, nl = 0;
        CK(hkp::d2h(&ns, e->er.n_spec, sizeof(int), e->stream));
        CK(hkp::d2h(&nl, e->er.n_slots, sizeof(int), e->stream));
        e->spec_h.resize(ns);
        CK(hkp::d2h(e->spec_h.data(), e->d.spec, (size_t)ns * sizeof(HkSpecialNode), e->stream));
        CK(hkp::d2h(e->spec_idx_h.data(), e->d.spec_idx, e->spec_idx_h.size() * sizeof(int), e->stream));
        e->n_contact_slots = nl;
    }
    e->contact_host_stale = false;
    return 0;
}

// fetch deletion log entries [del_seen, count) -> sorted (step, id), appended to deleted_all
static int fetch_deleted(hk_engine* e, std::vector<int64_t>* fresh) {
    if (e->dev_erosion) {
        int ovf = 0;
        CK(hkp::d2h(&ovf, e->er.overflow, sizeof(int), e->stream));
        if (ovf) return fail(e, HK_ERR_STATE, "contact surface grew beyond its device capacity");
    }
    int count = 0;
    CK(hkp::d2h(&count, e->d.del_count, sizeof(int), e->stream));
    if (count > e->d.del_cap) count = e->d.del_cap;
    if (count <= e->del_seen) return 0;
    std::vector<long long> ent(count - e->del_seen);
    CK(hkp::d2h(ent.data(), e->d.del_list + e->del_seen, ent.size() * sizeof(long long), e->stream));
    // already in the reference's order: the deletion pass appends ascending ids, steps are stream-ordered
    for (long long v : ent) {
        int64_t id = (int64_t)(v & 0xffffffffll) + 1;
        e->deleted_all.push_back(id);
        e->deleted_step.push_back((int64_t)(v >> 32));
        if (fresh) fresh->push_back(id);
    }
    e->del_seen = count;
    return 0;
}

// --------------------------------------------------------------------------------- NCCL, loaded at run time
// The engine owns its communicator.  libnccl is not a link-time dependency: the copy already loaded in the
// process (e.g. by the host framework) is used when there is one, else libnccl.so.2 from the loader path.
#ifndef HK_EMU
#include <dlfcn.h>
namespace {
struct HkNcclId { char internal[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(HkNcclId*) = nullptr;
    int (*CommInitRank)(void**, int, HkNcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
NcclApi g_nccl;
const int kNcclFloat64 = 8;          // ncclFloat64 (nccl.h)
const int kNcclInt64 = 4;            // ncclInt64
const int kNcclSum = 0;              // ncclSum

bool nccl_load() {
    if (g_nccl.lib) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_nccl.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) g_nccl.err = std::string("libnccl lacks ") + n; return p; };
    g_nccl.GetUniqueId = (int (*)(HkNcclId*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, HkNcclId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.Send || !g_nccl.Recv ||
        !g_nccl.AllGather || !g_nccl.AllReduce ||
        !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.GetErrorString)
        return false;
    g_nccl.lib = h;
    return true;
}
}  // namespace
#define NCK(call) do { int rn_ = (call); if (rn_) return fail(e, HK_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(rn_)); } while (0)
#endif

// own partial forces of all interface nodes -> d_halo_own, then one send block per neighbour
static int halo_pack_all(hk_engine* e) {
    if (e->halo.empty()) return 0;
    hk_launch_halo_pack(e->d, e->d_halo_list, e->n_halo_nodes, e->d_halo_own, e->use_Q0 ? e->d.Q0 : nullptr, e->stream);
    e->n_launch += 1;
    for (HaloNbr& h : e->halo) {
        if (!h.send) return fail(e, HK_ERR_STATE, "hk_halo_bind (or hk_comm_init) not called for every neighbour");
        hk_launch_halo_gather(e->d_halo_own, h.d_slots, (long long)h.slots.size(), h.send, e->stream);
        e->n_launch += 1;
    }
    return 0;
}

// received partials + this rank's own -> d.halo_recv, one holder at a time in ascending global-rank order (own first
// when the ranks are unknown): every holder of an interface node then forms the same sum bit for bit, whatever the
// number of holders
static int halo_total(hk_engine* e) {
    const HkDev& d = e->d;
    CK(hkp::dev_memset(d.halo_recv, 0, sizeof(double) * 3 * e->n_halo_nodes, e->stream));
    std::vector<size_t> order(e->halo.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return e->halo[a].rank < e->halo[b].rank; });
    bool own_done = false;
    auto add_own = [&]() {
        hk_launch_halo_accumulate(d, nullptr, e->n_halo_nodes, e->d_halo_own, e->stream);
        e->n_launch += 1;
        own_done = true;
    };
    for (size_t i : order) {
        HaloNbr& h = e->halo[i];
        if (!h.recv) return fail(e, HK_ERR_STATE, "hk_halo_bind (or hk_comm_init) not called for every neighbour");
        if (!own_done && (e->my_rank < 0 || h.rank > e->my_rank)) add_own();
        hk_launch_halo_accumulate(d, h.d_slots, (long long)h.slots.size(), h.recv, e->stream);
        e->n_launch += 1;
    }
    if (!own_done) add_own();
    return 0;
}

// ================================================================================= exported ABI
extern "C" {

int HKAPI(default_params)(hk_params* p) {
    if (!p) return HK_ERR_ARG;
    std::memset(p, 0, sizeof(*p));
    p->struct_size = (int32_t)sizeof(hk_params);
    p->contact_d_lim_factor = 0.3;
    p->contact_myu = 0.25;
    p->contact_kc_other = 1.0;
    p->contact_kc_self = 1.0;
    p->contact_cr_other = 0.0;
    p->contact_cr_self = 0.0;
    p->contact_ddiv_other = 1.1;
    p->contact_ddiv_self = 0.6;
    p->deterministic = 1;
    return HK_OK;
}

int HKAPI(create)(hk_engine** out, const hk_params* p) {
    if (!out || !p) return fail(nullptr, HK_ERR_ARG, "null argument");
    if (p->struct_size != (int32_t)sizeof(hk_params)) return fail(nullptr, HK_ERR_ARG, "hk_params size mismatch");
    if (!(p->d_time > 0)) return fail(nullptr, HK_ERR_ARG, "d_time must be > 0");
    if (p->triax_route != 0) return fail(nullptr, HK_ERR_UNSUPPORTED, "triax_route 1 (eigenvalue route) exists only in the oracle");
#ifndef HK_EMU
    int ndev = 0;
    cudaError_t rc = cudaGetDeviceCount(&ndev);
    if (rc != cudaSuccess || ndev == 0)
        return fail(nullptr, HK_ERR_NO_DEVICE, std::string("no CUDA device (") + cudaGetErrorString(rc) +
                                                   "): this engine has no CPU fallback");
    if (p->device < 0 || p->device >= ndev) return fail(nullptr, HK_ERR_ARG, "bad device ordinal");
    rc = cudaSetDevice(p->device);
    if (rc != cudaSuccess) return fail(nullptr, HK_ERR_CUDA, cudaGetErrorString(rc));
#endif
    hk_engine* e = new hk_engine();
    e->prm = *p;
    std::memset(&e->d, 0, sizeof(e->d));
#ifndef HK_EMU
    rc = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (rc != cudaSuccess) { delete e; return fail(nullptr, HK_ERR_CUDA, cudaGetErrorString(rc)); }
    e->own_stream = true;
    if (const char* g = getenv("HK_STEP_GRAPH")) e->graph_off = atoi(g) == 0;      // HK_STEP_GRAPH=0: plain launches only
#endif
    *out = e;
    return HK_OK;
}

int HKAPI(destroy)(hk_engine* e) {
    if (!e) return HK_OK;
#ifndef HK_EMU
    cudaSetDevice(e->prm.device);
#endif
    prof_collect(e);
    hkp::sync(e->stream);
    for (void* p : e->allocs) hkp::dev_free(p);
#ifndef HK_EMU
    if (e->step_graph) cudaGraphExecDestroy(e->step_graph);
    if (e->comm) {
        cudaStreamSynchronize(e->comm_stream);
        g_nccl.CommDestroy(e->comm);
        cudaStreamDestroy(e->comm_stream);
        cudaEventDestroy(e->ev_pack);
        cudaEventDestroy(e->ev_comm);
    }
    if (e->copy_stream) {
        cudaStreamDestroy(e->copy_stream);
        for (int b = 0; b < 2; ++b) { cudaEventDestroy(e->ev_ready[b]); cudaEventDestroy(e->ev_free[b]); }
    }
    if (e->own_stream) cudaStreamDestroy(e->stream);
#endif
    delete e;
    return HK_OK;
}

const char* HKAPI(last_error)(const hk_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int HKAPI(set_mesh)(hk_engine* e, int64_t nNode, int64_t nElement, const double* coordmat, const int64_t* elementmat,
                    const int64_t* element_material, const int64_t* element_instance, const double* diag_M) {
    if (!e || !coordmat || !elementmat || !element_material || !diag_M) return fail(e, HK_ERR_ARG, "null argument");
    if (e->finalized) return fail(e, HK_ERR_STATE, "already finalised");
    if (nNode <= 0 || nElement <= 0 || nElement >= (1ll << 28) || nNode >= (1ll << 31) / 6)
        return fail(e, HK_ERR_ARG, "mesh size out of range (nElement < 2^28)");
    e->nNode = nNode;
    e->nElement = nElement;
    e->coordmat.assign(coordmat, coordmat + 3 * nNode);
    e->mass.resize(nNode);
    for (int64_t n = 0; n < nNode; ++n) {
        const double m = diag_M[3 * n];
        if (diag_M[3 * n + 1] != m || diag_M[3 * n + 2] != m)
            return fail(e, HK_ERR_ARG, "diag_M must hold the same mass on the 3 dofs of a node (J2:205-215)");
        e->mass[n] = m;
    }
    e->conn.resize(8 * nElement);
    for (int64_t i = 0; i < 8 * nElement; ++i) {
        int64_t v = elementmat[i];
        if (v < 1 || v > nNode) return fail(e, HK_ERR_ARG, "elementmat entry out of range");
        e->conn[i] = (int)(v - 1);
    }
    e->emat.resize(nElement);
    for (int64_t i = 0; i < nElement; ++i) e->emat[i] = (int)(element_material[i] - 1);
    e->einst.assign(nElement, 1);
    if (element_instance) e->einst.assign(element_instance, element_instance + nElement);
    return HK_OK;
}

int HKAPI(add_material)(hk_engine* e, double young, double poisson, double density, int64_t npp, const double* plastic,
                        const double* Hd, int64_t nd, const double* ductile) {
    if (!e) return HK_ERR_ARG;
    if (e->finalized) return fail(e, HK_ERR_STATE, "already finalised");
    if (npp == 1) return fail(e, HK_ERR_ARG, "*Plastic table needs >= 2 rows (the reference indexes Hd[1])");
    if (npp < 0 || nd < 0 || (npp > 0 && (!plastic || !Hd)) || (nd > 0 && !ductile))
        return fail(e, HK_ERR_ARG, "material table pointer is NULL (or a negative row count)");
    if (npp > HK_MAX_TABLE || nd > HK_MAX_TABLE) return fail(e, HK_ERR_UNSUPPORTED, "material table longer than HK_MAX_TABLE");
    for (int64_t r = 1; r < npp; ++r)
        if (!(plastic[npp + r] > plastic[npp + r - 1]))
            return fail(e, HK_ERR_UNSUPPORTED, "*Plastic table: equivalent plastic strain column must be increasing");
    MaterialH m;
    m.young = young; m.poisson = poisson; m.density = density; m.npp = npp; m.nd = nd;
    if (npp > 0) { m.plastic.assign(plastic, plastic + 2 * npp); m.Hd.assign(Hd, Hd + npp - 1); }
    if (nd > 0) m.ductile.assign(ductile, ductile + 3 * nd);
    e->materials.push_back(m);
    return HK_OK;
}

int HKAPI(add_bc)(hk_engine* e, int64_t n_lists, const int64_t* dof_ptr, const int64_t* dofs, const double* values,
                  int64_t n_amp, const double* amp_time, const double* amp_value) {
    if (!e) return HK_ERR_ARG;
    if (e->finalized) return fail(e, HK_ERR_STATE, "already finalised");
    if (n_lists < 0 || (n_lists > 0 && (!dof_ptr || !values))) return fail(e, HK_ERR_ARG, "hk_add_bc: null argument");
    BCH b;
    for (int64_t j = 0; j < n_lists; ++j) {
        if (dof_ptr[j + 1] < dof_ptr[j] || (dof_ptr[j + 1] > dof_ptr[j] && !dofs)) return fail(e, HK_ERR_ARG, "hk_add_bc: bad dof_ptr");
        b.dof.emplace_back(dofs + dof_ptr[j], dofs + dof_ptr[j + 1]);
        b.value.push_back(values[j]);
    }
    if (n_amp > 0) {
        if (n_amp < 2) return fail(e, HK_ERR_ARG, "amplitude table needs >= 2 points");
        b.a_t.assign(amp_time, amp_time + n_amp);
        b.a_v.assign(amp_value, amp_value + n_amp);
    }
    e->bcs.push_back(b);
    return HK_OK;
}

int HKAPI(add_ic)(hk_engine* e, int64_t n_lists, const int64_t* dof_ptr, const int64_t* dofs, const double* values) {
    if (!e) return HK_ERR_ARG;
    if (e->finalized) return fail(e, HK_ERR_STATE, "already finalised");
    if (n_lists < 0 || (n_lists > 0 && (!dof_ptr || !values))) return fail(e, HK_ERR_ARG, "hk_add_ic: null argument");
    ICH b;
    for (int64_t j = 0; j < n_lists; ++j) {
        b.dof.emplace_back(dofs + dof_ptr[j], dofs + dof_ptr[j + 1]);
        b.value.push_back(values[j]);
    }
    e->ics.push_back(b);
    return HK_OK;
}

int HKAPI(add_instance)(hk_engine* e, int64_t node_offset, int64_t nNode, int64_t element_offset, int64_t nElement,
                        const int64_t* surfaces, const int64_t* surfaces_eleid) {
    if (!e) return HK_ERR_ARG;
    if (e->finalized) return fail(e, HK_ERR_STATE, "already finalised");
    if (node_offset < 0 || nNode < 0 || element_offset < 0 || nElement < 0) return fail(e, HK_ERR_ARG, "hk_add_instance: negative range");
    InstanceH I;
    I.node_offset = node_offset; I.nNode = nNode; I.element_offset = element_offset; I.nElement = nElement;
    if (surfaces && surfaces_eleid) {
        for (int64_t k = 0; k < 24 * nElement; ++k)
            if (surfaces[k] < 1 || surfaces[k] > nNode) return fail(e, HK_ERR_ARG, "hk_add_instance: face node outside the instance");
        for (int64_t k = 0; k < 6 * nElement; ++k)
            if (surfaces_eleid[k] < 1 || surfaces_eleid[k] > nElement) return fail(e, HK_ERR_ARG, "hk_add_instance: face element outside the instance");
        I.surfaces.assign(surfaces, surfaces + 24 * nElement);
        I.eleid.assign(surfaces_eleid, surfaces_eleid + 6 * nElement);
    }
    e->instances.push_back(std::move(I));
    return HK_OK;
}

int HKAPI(add_contact_pair)(hk_engine* e, int64_t i_instance, int64_t j_instance, int64_t nn_i, const int64_t* c_nodes_i,
                            int64_t nn_j, const int64_t* c_nodes_j, int64_t nTri, const int64_t* c_triangles,
                            const int64_t* c_triangles_eleid, double young) {
    if (!e) return HK_ERR_ARG;
    if (e->finalized) return fail(e, HK_ERR_STATE, "already finalised");
    if (e->nNode == 0) return fail(e, HK_ERR_STATE, "hk_set_mesh must precede hk_add_contact_pair");
    if (nn_i < 0 || nn_j < 0 || nTri < 0 || (nn_i > 0 && !c_nodes_i) || (nn_j > 0 && !c_nodes_j) ||
        (nTri > 0 && (!c_triangles || !c_triangles_eleid)))
        return fail(e, HK_ERR_ARG, "hk_add_contact_pair: null argument");
    for (int64_t k = 0; k < nn_i; ++k) if (c_nodes_i[k] < 1 || c_nodes_i[k] > e->nNode) return fail(e, HK_ERR_ARG, "c_nodes_i entry out of range");
    for (int64_t k = 0; k < nn_j; ++k) if (c_nodes_j[k] < 1 || c_nodes_j[k] > e->nNode) return fail(e, HK_ERR_ARG, "c_nodes_j entry out of range");
    for (int64_t k = 0; k < 3 * nTri; ++k) if (c_triangles[k] < 1 || c_triangles[k] > e->nNode) return fail(e, HK_ERR_ARG, "c_triangles entry out of range");
    for (int64_t k = 0; k < nTri; ++k) if (c_triangles_eleid[k] < 1 || c_triangles_eleid[k] > e->nElement) return fail(e, HK_ERR_ARG, "c_triangles_eleid entry out of range");
    PairH p;
    p.i_instance = i_instance; p.j_instance = j_instance; p.young = young;
    std::memset(&p.dev, 0, sizeof(p.dev));
    for (int64_t k = 0; k < nn_i; ++k) { p.nodes_i.push_back((int)(c_nodes_i[k] - 1)); p.set_i.insert((int)(c_nodes_i[k] - 1)); }
    for (int64_t k = 0; k < nn_j; ++k) { p.nodes_j.push_back((int)(c_nodes_j[k] - 1)); p.set_j.insert((int)(c_nodes_j[k] - 1)); }
    for (int64_t k = 0; k < nTri; ++k) {
        p.t0.push_back((int)(c_triangles[k] - 1));
        p.t1.push_back((int)(c_triangles[k + nTri] - 1));
        p.t2.push_back((int)(c_triangles[k + 2 * nTri] - 1));
        p.tele.push_back((int)(c_triangles_eleid[k] - 1));
    }
    e->pairs.push_back(std::move(p));
    return HK_OK;
}

// Contact set-up on the device (A12; hk_setup.cu): replaces the host's get_element_face / get_surface_triangle and the
// hk_add_instance / hk_add_contact_pair calls built from them.
int HKAPI(build_contact)(hk_engine* e, int64_t n_inst, const int64_t* node_offset, const int64_t* nNode,
                         const int64_t* element_offset, const int64_t* nElement, const double* young, int64_t n_cp,
                         const int64_t* cp_inst1, const int64_t* cp_inst2, const int64_t* cp_ptr1, const int64_t* cp_elems1,
                         const int64_t* cp_ptr2, const int64_t* cp_elems2) {
    if (!e) return HK_ERR_ARG;
    if (e->finalized) return fail(e, HK_ERR_STATE, "already finalised");
    if (e->nNode == 0) return fail(e, HK_ERR_STATE, "hk_set_mesh must precede hk_build_contact");
    if (!e->instances.empty() || !e->pairs.empty()) return fail(e, HK_ERR_STATE, "hk_build_contact replaces hk_add_instance / hk_add_contact_pair");
    if (n_inst < 1 || !node_offset || !nNode || !element_offset || !nElement || !young) return fail(e, HK_ERR_ARG, "null argument");
    if (n_cp < 0 || (n_cp > 0 && (!cp_inst1 || !cp_inst2))) return fail(e, HK_ERR_ARG, "null argument");
    for (int64_t i = 0; i < n_inst; ++i)
        if (node_offset[i] < 0 || nNode[i] < 0 || node_offset[i] + nNode[i] > e->nNode || element_offset[i] < 0 || nElement[i] < 0 ||
            element_offset[i] + nElement[i] > e->nElement)
            return fail(e, HK_ERR_ARG, "instance range outside the mesh");
    HK_DEVICE(e);
    // the pair list (J2:272-314 for *Contact Inclusions, ALL EXTERIOR; else the given *Contact Pair list)
    struct CPH { int64_t i1, i2; std::vector<int64_t> el1, el2; };
    std::vector<CPH> cps;
    if (n_cp == 0) {
        if (n_inst > 1) {
            for (int64_t i = 1; i <= n_inst; ++i)
                for (int64_t j = (e->prm.contact_flag == 2 ? i : i + 1); j <= n_inst; ++j) cps.push_back({i, j, {}, {}});
        } else {
            cps.push_back({1, 1, {}, {}});
        }
    } else {
        for (int64_t k = 0; k < n_cp; ++k) {
            if (cp_inst1[k] < 1 || cp_inst1[k] > n_inst || cp_inst2[k] < 1 || cp_inst2[k] > n_inst) return fail(e, HK_ERR_ARG, "contact pair instance out of range");
            CPH c{cp_inst1[k], cp_inst2[k], {}, {}};
            if (cp_ptr1 && cp_elems1) c.el1.assign(cp_elems1 + cp_ptr1[k], cp_elems1 + cp_ptr1[k + 1]);
            if (cp_ptr2 && cp_elems2) c.el2.assign(cp_elems2 + cp_ptr2[k], cp_elems2 + cp_ptr2[k + 1]);
            cps.push_back(std::move(c));
        }
    }
    // temporary device copies of the mesh (hk_finalize builds the resident ones)
    const int64_t nE = e->nElement, nN = e->nNode;
    int* d_conn = nullptr;
    double* d_X = nullptr;
    {
        std::vector<int> soa((size_t)8 * nE);
        for (int64_t el = 0; el < nE; ++el)
            for (int a = 0; a < 8; ++a) soa[(size_t)a * nE + el] = e->conn[8 * el + a];
        int rc;
        if ((rc = dalloc(e, &d_conn, soa.size()))) return rc;
        if ((rc = dalloc(e, &d_X, (size_t)3 * nN))) return rc;
        if ((rc = upload(e, d_conn, soa))) return rc;
        if ((rc = upload(e, d_X, e->coordmat))) return rc;
    }
    std::vector<std::vector<int>> exterior(n_inst);
    e->instances.clear();
    for (int64_t i = 0; i < n_inst; ++i) {
        InstanceH I;
        I.node_offset = node_offset[i]; I.nNode = nNode[i]; I.element_offset = element_offset[i]; I.nElement = nElement[i];
        std::vector<int> surf;
        const int rc = hk_setup_instance_faces(d_conn, nE, d_X, I.node_offset, I.element_offset, I.nElement, e->stream, surf,
                                               I.twin, exterior[i]);
        if (rc) { dfree(e, d_conn); dfree(e, d_X); return cuda_fail(e, rc, "hk_setup_instance_faces"); }
        I.surfaces.assign(surf.begin(), surf.end());
        I.eleid.resize((size_t)6 * I.nElement);
        for (int64_t f = 0; f < 6 * I.nElement; ++f) I.eleid[f] = f / 6 + 1;
        e->instances.push_back(std::move(I));
    }
    dfree(e, d_conn);
    dfree(e, d_X);
    // get_surface_triangle's tail (J2:2094-2159) on the exterior faces: optional *Surface element filter, two triangles
    // per face, sorted unique nodes — all part-local 1-based
    struct Surf { std::vector<int64_t> tri, te, nodes; };
    auto surface_of = [&](int64_t inst, const std::vector<int64_t>& subset, Surf& out) {
        const InstanceH& I = e->instances[inst - 1];
        const int64_t F = 6 * I.nElement;
        std::vector<char> keep;
        if (!subset.empty() && (int64_t)subset.size() != I.nElement) {
            keep.assign(I.nElement + 1, 0);
            for (int64_t el : subset) if (el >= 1 && el <= I.nElement) keep[el] = 1;
        }
        for (int f : exterior[inst - 1]) {
            const int64_t el = f / 6 + 1;
            if (!keep.empty() && !keep[el]) continue;
            const int64_t s0 = I.surfaces[f], s1 = I.surfaces[f + F], s2 = I.surfaces[f + 2 * F], s3 = I.surfaces[f + 3 * F];
            out.tri.insert(out.tri.end(), {s0, s1, s2, s2, s3, s0});
            out.te.push_back(el); out.te.push_back(el);
        }
        out.nodes = out.tri;
        std::sort(out.nodes.begin(), out.nodes.end());
        out.nodes.erase(std::unique(out.nodes.begin(), out.nodes.end()), out.nodes.end());
    };
    for (const CPH& c : cps) {
        Surf s1, s2;
        surface_of(c.i1, c.el1, s1);
        surface_of(c.i2, c.el2, s2);
        const int n_dir = c.i1 == c.i2 ? 1 : 2;                   // J2:339-354
        for (int dir = 0; dir < n_dir; ++dir) {
            const int64_t ii = dir == 0 ? c.i1 : c.i2, jj = dir == 0 ? c.i2 : c.i1;
            const Surf& si = dir == 0 ? s1 : s2;
            const Surf& sj = dir == 0 ? s2 : s1;
            const InstanceH& Ii = e->instances[ii - 1];
            const InstanceH& Ij = e->instances[jj - 1];
            PairH p;
            p.i_instance = ii; p.j_instance = jj; p.young = young[jj - 1];          // J2:372
            std::memset(&p.dev, 0, sizeof(p.dev));
            for (int64_t n : si.nodes) { const int g = (int)(n + Ii.node_offset - 1); p.nodes_i.push_back(g); p.set_i.insert(g); }
            for (int64_t n : sj.nodes) { const int g = (int)(n + Ij.node_offset - 1); p.nodes_j.push_back(g); p.set_j.insert(g); }
            for (size_t t = 0; t < sj.te.size(); ++t) {
                p.t0.push_back((int)(sj.tri[3 * t] + Ij.node_offset - 1));
                p.t1.push_back((int)(sj.tri[3 * t + 1] + Ij.node_offset - 1));
                p.t2.push_back((int)(sj.tri[3 * t + 2] + Ij.node_offset - 1));
                p.tele.push_back((int)(sj.te[t] + Ij.element_offset - 1));
            }
            e->pairs.push_back(std::move(p));
        }
    }
    return HK_OK;
}

int HKAPI(finalize)(hk_engine* e) {
    if (!e) return HK_ERR_ARG;
    if (e->finalized) return fail(e, HK_ERR_STATE, "already finalised");
    if (e->nNode == 0) return fail(e, HK_ERR_STATE, "hk_set_mesh not called");
    HK_DEVICE(e);
    const int64_t nN = e->nNode, nE = e->nElement;
    HkDev& d = e->d;
    d.variant = hk_element_variant_from_env();
    d.n_sm = 1;
#ifndef HK_EMU
    CK(cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, e->prm.device));
#endif
    const int64_t tile = hk_element_tile(d.variant);
    const int64_t nEp = (nE + tile - 1) / tile * tile;   // whole tiles for the element kernel
    d.nNode = nN; d.nElement = nE; d.nEp = nEp;
    d.element_mode = e->prm.element_mode;
    const double dt = e->prm.d_time;
    e->dt2 = dt * dt;                 // d_time^2   (J2:564)
    e->dt2p = std::pow(dt, 2.0);      // d_time^2.0 (J2:564)
    int rc;

    // ---- materials
    if (e->materials.empty()) return fail(e, HK_ERR_STATE, "no material");
    std::vector<HkMaterialDev> md(e->materials.size());
    for (size_t i = 0; i < md.size(); ++i) {
        const MaterialH& m = e->materials[i];
        HkMaterialDev& D = md[i];
        std::memset(&D, 0, sizeof(D));
        D.young = m.young; D.poisson = m.poisson;
        D.G = m.young / 2. / (1.0 + m.poisson);                                  // J2:146
        const double d1 = (1.0 - m.poisson), d2 = m.poisson, d3 = (1.0 - 2.0 * m.poisson) / 2.0;
        const double f = m.young / (1.0 + m.poisson) / (1.0 - 2.0 * m.poisson);  // J2:153
        D.D11 = f * d1; D.D12 = f * d2; D.D44 = f * d3;
        D.npp = (int)m.npp; D.nd = (int)m.nd;
        for (int64_t r = 0; r < m.npp; ++r) { D.plastic_s[r] = m.plastic[r]; D.plastic_e[r] = m.plastic[r + m.npp]; }
        for (int64_t r = 0; r + 1 < m.npp; ++r) D.Hd[r] = m.Hd[r];
        for (int64_t r = 0; r < m.nd; ++r) { D.duct_e[r] = m.ductile[r]; D.duct_t[r] = m.ductile[r + m.nd]; }
        if (m.nd > 0) e->any_ductile = true;
    }
    for (int64_t i = 0; i < nE; ++i)
        if (e->emat[i] < 0 || e->emat[i] >= (int)md.size()) return fail(e, HK_ERR_ARG, "element_material out of range");
    d.n_mat = (int)md.size();
    if ((rc = dalloc(e, &d.mats, md.size()))) return rc;
    if ((rc = upload(e, d.mats, md))) return rc;
    {
        double P[8][3][8];
        cal_Pusai_hexa(P);
        CK(hk_upload_pusai(&P[0][0][0]));
    }

    // ---- nodes
    if ((rc = dalloc(e, &d.X, (size_t)3 * nN))) return rc;
    if ((rc = dalloc(e, &d.u, (size_t)3 * nN))) return rc;
    if ((rc = dalloc(e, &d.u_pre, (size_t)3 * nN))) return rc;
    if ((rc = dalloc(e, &d.velo, (size_t)3 * nN))) return rc;
    if ((rc = dalloc(e, &d.rec, (size_t)6 * nN))) return rc;
    if ((rc = dalloc(e, &d.mass, (size_t)nN))) return rc;
    if ((rc = dalloc(e, &d.Q0, (size_t)3 * nN))) return rc;
    if ((rc = upload(e, d.X, e->coordmat))) return rc;
    if ((rc = upload(e, d.mass, e->mass))) return rc;
    {
        std::vector<double> up(3 * nN, 0.0), ve(3 * nN, 0.0), rec(6 * nN, 0.0);
        for (const ICH& ic : e->ics)                                             // J2:233-239
            for (size_t j = 0; j < ic.dof.size(); ++j)
                for (int64_t dof : ic.dof[j]) {
                    if (dof < 1 || dof > 3 * nN) return fail(e, HK_ERR_ARG, "IC dof out of range");
                    up[dof - 1] = -ic.value[j] * dt;
                    ve[dof - 1] = ic.value[j];
                }
        for (int64_t n = 0; n < nN; ++n)
            for (int c = 0; c < 3; ++c) rec[6 * n + c] = e->coordmat[3 * n + c];   // position = coordmat, J2:222
        if ((rc = upload(e, d.u_pre, up))) return rc;
        if ((rc = upload(e, d.velo, ve))) return rc;
        if ((rc = upload(e, d.rec, rec))) return rc;
        CK(hkp::dev_memset(d.u, 0, sizeof(double) * 3 * nN, e->stream));
        CK(hkp::dev_memset(d.Q0, 0, sizeof(double) * 3 * nN, e->stream));
    }

    // ---- node -> element table (ELL, ascending element order = the reference's scatter order J2:669-675)
    {
        std::vector<int> cnt(nN, 0);
        for (int64_t i = 0; i < 8 * nE; ++i) cnt[e->conn[i]]++;
        int w = 0;
        for (int64_t n = 0; n < nN; ++n) w = std::max(w, cnt[n]);
        d.ell_width = w;
        std::vector<int> ell((size_t)w * nN, -1);
        std::fill(cnt.begin(), cnt.end(), 0);
        for (int64_t el = 0; el < nE; ++el)
            for (int a = 0; a < 8; ++a) {
                const int n = e->conn[8 * el + a];
                ell[(size_t)cnt[n] * nN + n] = (int)(el * 8 + a);
                cnt[n]++;
            }
        if ((rc = dalloc(e, &d.ell, ell.size()))) return rc;
        if ((rc = upload(e, d.ell, ell))) return rc;
    }

    // ---- boundary conditions -> special nodes
    e->spec_idx_h.assign(nN, -1);
    {
        std::vector<double> bc_value, amp_time, amp_value;
        std::vector<int> bc_amp;
        std::vector<HkAmpTable> amp_tab;
        for (const BCH& b : e->bcs) {
            int aid = -1;
            if (!b.a_t.empty()) {
                aid = (int)amp_tab.size();
                HkAmpTable t;
                t.n = (int)b.a_t.size();
                t.offset = (int)amp_time.size();
                amp_tab.push_back(t);
                amp_time.insert(amp_time.end(), b.a_t.begin(), b.a_t.end());
                amp_value.insert(amp_value.end(), b.a_v.begin(), b.a_v.end());
            }
            for (size_t j = 0; j < b.dof.size(); ++j) {
                const int entry = (int)bc_value.size();
                bc_value.push_back(b.value[j]);
                bc_amp.push_back(aid);
                for (int64_t dof : b.dof[j]) {
                    if (dof < 1 || dof > 3 * nN) return fail(e, HK_ERR_ARG, "BC dof out of range");
                    const int node = (int)((dof - 1) / 3), c = (int)((dof - 1) % 3);
                    const int si = spec_of(e, node);
                    e->spec_h[si].bc_entry[c] = entry;                    // later BCs override earlier ones
                }
            }
        }
        if ((rc = dalloc(e, &d.bc_value, bc_value.size()))) return rc;
        if ((rc = dalloc(e, &d.bc_amp, bc_amp.size()))) return rc;
        if ((rc = dalloc(e, &d.amp_tab, amp_tab.size()))) return rc;
        if ((rc = dalloc(e, &d.amp_time, amp_time.size()))) return rc;
        if ((rc = dalloc(e, &d.amp_value, amp_value.size()))) return rc;
        if ((rc = upload(e, d.bc_value, bc_value))) return rc;
        if ((rc = upload(e, d.bc_amp, bc_amp))) return rc;
        if ((rc = upload(e, d.amp_tab, amp_tab))) return rc;
        if ((rc = upload(e, d.amp_time, amp_time))) return rc;
        if ((rc = upload(e, d.amp_value, amp_value))) return rc;
    }

    // ---- elements
    {
        std::vector<int> conn_soa((size_t)8 * nEp, 0);
        for (int64_t el = 0; el < nE; ++el)
            for (int a = 0; a < 8; ++a) conn_soa[(size_t)a * nEp + el] = e->conn[8 * el + a];
        if ((rc = dalloc(e, &d.conn, conn_soa.size()))) return rc;
        if ((rc = upload(e, d.conn, conn_soa))) return rc;
        std::vector<unsigned char> fl(nEp, 2);
        std::vector<unsigned short> mt(nEp, 0);
        for (int64_t el = 0; el < nE; ++el) { fl[el] = 1; mt[el] = (unsigned short)e->emat[el]; }
        if ((rc = dalloc(e, &d.flag, (size_t)nEp))) return rc;
        if ((rc = dalloc(e, &d.mat, (size_t)nEp))) return rc;
        if ((rc = upload(e, d.flag, fl))) return rc;
        if ((rc = upload(e, d.mat, mt))) return rc;
    }
    d.TL = (int)tile;
    if ((rc = dalloc(e, &d.ips, (size_t)112 * nEp))) return rc;
    if ((rc = dalloc(e, &d.triax, (size_t)8 * nEp))) return rc;
    if ((rc = dalloc(e, &d.Qe, (size_t)24 * nEp))) return rc;
    CK(hkp::dev_memset(d.ips, 0, sizeof(double) * 112 * nEp, e->stream));
    CK(hkp::dev_memset(d.triax, 0, sizeof(double) * 8 * nEp, e->stream));
    CK(hkp::dev_memset(d.Qe, 0, sizeof(double) * 24 * nEp, e->stream));
    {   // integ_yield_stress = plastic[1,1] of the element's material (J2:456-465)
        const HkDev dd = d;
        hk_parallel_for(nE * 8, e->stream, HK_LAMBDA(long long i) {
            const long long el = i % dd.nElement, k = i / dd.nElement;
            const HkMaterialDev& M = dd.mats[dd.mat[el]];
            if (M.npp > 0) dd.ips[hk_ip(dd, 13, (int)k, el)] = M.plastic_s[0];
        });
    }
    d.del_cap = (int)nE;
    if ((rc = dalloc(e, &d.del_count, (size_t)1))) return rc;
    if ((rc = dalloc(e, &d.del_list, (size_t)nE))) return rc;
    if ((rc = dalloc(e, &d.del_block, (size_t)((nE + 1023) / 1024)))) return rc;
    if ((rc = dalloc(e, &d.del_fresh, (size_t)1))) return rc;
    CK(hkp::dev_memset(d.del_fresh, 0, sizeof(int), e->stream));
    if ((rc = dalloc(e, &d.counters, (size_t)8))) return rc;
    CK(hkp::dev_memset(d.del_count, 0, sizeof(int), e->stream));
    CK(hkp::dev_memset(d.counters, 0, 8 * sizeof(unsigned long long), e->stream));

    // ---- contact
    if ((rc = dalloc(e, &d.spec_idx, (size_t)nN))) return rc;
    if (e->prm.contact_flag >= 1) {
        for (PairH& p : e->pairs) {
            for (int n : p.nodes_i) ensure_contact_slot(e, n, nullptr);
            for (int n : p.nodes_j) ensure_contact_slot(e, n, nullptr);
            for (size_t k = 0; k < p.t0.size(); ++k) {     // triangle vertices are members of c_nodes_j by construction
                ensure_contact_slot(e, p.t0[k], nullptr);
                ensure_contact_slot(e, p.t1[k], nullptr);
                ensure_contact_slot(e, p.t2[k], nullptr);
            }
            if ((rc = pair_upload(e, p))) return rc;
        }
        HkContactParams& cp = e->cp;
        cp.d_lim = e->prm.element_min_size * e->prm.contact_d_lim_factor;        // J2:2254
        cp.myu = e->prm.contact_myu;
        cp.kc_o = e->prm.contact_kc_other; cp.kc_s = e->prm.contact_kc_self;
        cp.cr_o = e->prm.contact_cr_other; cp.cr_s = e->prm.contact_cr_self;
        cp.ddiv_o = e->prm.element_max_size * e->prm.contact_ddiv_other;          // J2:2331
        cp.ddiv_s = e->prm.element_max_size * e->prm.contact_ddiv_self;           // J2:2333
        cp.d_time = dt;
        cp.clamp = 0; cp.dnode = nullptr; cp.dnode_pre = nullptr; cp.dmax = nullptr;
        double ymax = 0;
        for (const PairH& p : e->pairs) ymax = std::max(ymax, p.young);
        const double nominal = ymax * e->prm.element_max_size * cp.d_lim * std::max(cp.kc_o, cp.kc_s);
        int ex = 0;
        if (nominal > 0) std::frexp(nominal, &ex);
        cp.lsb_exp = ex - 90;      // 2^36 of headroom above the nominal force E*eMax*d_lim, 90 bits below it
    }
    // ---- halo (multi-GPU): dense slot per interface node; neighbours contribute in list order
    if (!e->halo.empty()) {
        std::vector<int> slot_of(nN, -1);
        for (HaloNbr& h : e->halo) {
            h.slots.resize(h.nodes.size());
            for (size_t i = 0; i < h.nodes.size(); ++i) {
                const int nd = h.nodes[i];
                if (nd < 0 || nd >= nN) return fail(e, HK_ERR_ARG, "halo node out of range");
                if (slot_of[nd] < 0) {
                    slot_of[nd] = e->n_halo_nodes++;
                    e->spec_h[spec_of(e, nd)].halo_slot = slot_of[nd];
                }
                h.slots[i] = slot_of[nd];
            }
            if ((rc = dalloc(e, &h.d_nodes, h.nodes.size()))) return rc;
            if ((rc = dalloc(e, &h.d_slots, h.slots.size()))) return rc;
            if ((rc = upload(e, h.d_nodes, h.nodes))) return rc;
            if ((rc = upload(e, h.d_slots, h.slots))) return rc;
        }
        {
            std::vector<int> list(e->n_halo_nodes, 0);
            for (int64_t nd = 0; nd < nN; ++nd) if (slot_of[nd] >= 0) list[slot_of[nd]] = (int)nd;
            if ((rc = dalloc(e, &e->d_halo_list, list.size()))) return rc;
            if ((rc = upload(e, e->d_halo_list, list))) return rc;
        }
        if ((rc = dalloc(e, &e->d_halo_own, (size_t)3 * e->n_halo_nodes))) return rc;
        if ((rc = dalloc(e, &d.halo_recv, (size_t)3 * e->n_halo_nodes))) return rc;
        CK(hkp::dev_memset(d.halo_recv, 0, sizeof(double) * 3 * e->n_halo_nodes, e->stream));
    }
    if ((rc = spec_upload(e, nullptr))) return rc;
    CK(hkp::sync(e->stream));
    CK(hkp::last_error());
    e->finalized = true;
    e->velo_current = true;
    e->triax_current = true;
    return HK_OK;
}

// contact pass of one step (A11 + the accumulator reset of A2): all ordered pairs on the engine's stream
static int contact_pass(hk_engine* e) {
    const HkDev& d = e->d;
    prof_begin(e, 0);
    if (e->dev_erosion) hk_launch_cacc_zero(d, e->er.n_slots, e->er.slot_cap, e->stream);
    else CK(hkp::dev_memset(d.cacc, 0, (size_t)e->n_contact_slots * 6 * sizeof(unsigned long long), e->stream));
    if (e->prm.contact_dmax_clamp) {              // v0.0.1 clamp: d_node .= 0 (J1:492); d_node_pre / d_max are last step's
        if (e->dnode_cap < e->cacc_cap) {
            dfree(e, e->d_dnode);
            e->d_dnode = nullptr;
            int rc = dalloc(e, &e->d_dnode, 2 * e->cacc_cap);
            if (rc) return rc;
            e->dnode_cap = e->cacc_cap;
            CK(hkp::dev_memset(e->d_dnode, 0, 2 * e->dnode_cap * sizeof(double), e->stream));
        }
        if (!e->d_dmax) {
            int rc = dalloc(e, &e->d_dmax, (size_t)2);
            if (rc) return rc;
            CK(hkp::dev_memset(e->d_dmax, 0, 2 * sizeof(unsigned long long), e->stream));      // d_max = 0.0, J1:413
        }
        const int cur = e->clamp_cur;
        e->cp.clamp = 1;
        e->cp.dnode = e->d_dnode + (size_t)cur * e->dnode_cap;
        e->cp.dnode_pre = e->d_dnode + (size_t)(1 - cur) * e->dnode_cap;
        e->cp.dmax = e->d_dmax + cur;
        CK(hkp::dev_memset(e->cp.dnode, 0, e->dnode_cap * sizeof(double), e->stream));
    }
    for (PairH& p : e->pairs) { hk_launch_contact(d, p.dev, e->cp, e->stream); e->n_launch += 5; }   // reset, bbox, cells, cull, narrow
    prof_end(e);
    return 0;
}

// enqueue the kernels of steps t_first .. t_first+n_steps-1 on the engine's stream (no host synchronisation
// unless contact surfaces may change: exposed faces must be in place before the next contact pass).
// phase 0: whole step; phase 1: everything that does not need the halo (contact + nodal update of non-interface
// nodes); phase 2: the rest (received partials, interface nodes, element kernel).
static int enqueue_steps(hk_engine* e, int64_t t_first, int64_t n_steps, bool frame_at_end, int phase);

#ifndef HK_EMU
// Captures the launches of one ordinary step (whatever enqueue_steps issues for it: contact pass, nodal update, element
// kernel, deletion pass) into a graph whose kernels read the step number from *d_step, plus a last node that advances it.
static int step_graph_capture(hk_engine* e) {
    if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
    if (!e->d_step) { int rc = dalloc(e, &e->d_step, (size_t)1); if (rc) return rc; }
    const long long launches0 = e->n_launch, steps0 = e->n_steps;
    const bool velo0 = e->velo_current, triax0 = e->triax_current, stale0 = e->contact_host_stale;
    e->d.t_dev = e->d_step;
    cudaGraph_t g = nullptr;
    int rc = 0;
    cudaError_t ce = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
    if (ce == cudaSuccess) {
        e->capturing = true;
        rc = enqueue_steps(e, 0, 1, false, 0);
        e->capturing = false;
        hk_launch_step_advance(e->d_step, e->stream);
        ce = cudaStreamEndCapture(e->stream, &g);
    }
    e->d.t_dev = nullptr;
    e->graph_launches = e->n_launch - launches0 + 1;
    e->n_launch = launches0; e->n_steps = steps0;                  // nothing ran
    e->velo_current = velo0; e->triax_current = triax0; e->contact_host_stale = stale0;
    if (rc == 0 && ce == cudaSuccess && g) ce = cudaGraphInstantiate(&e->step_graph, g, 0);
    if (g) cudaGraphDestroy(g);
    if (rc || ce != cudaSuccess || !e->step_graph) {                // keep running with plain launches (a real error
        cudaGetLastError();                                         // will show up again there, with its message)
        e->step_graph = nullptr;
        e->graph_off = true;
        return 0;
    }
    e->graph_dirty = false;
    return 0;
}
#endif

static int enqueue_steps(hk_engine* e, int64_t t_first, int64_t n_steps, bool frame_at_end, int phase) {
    if (phase != 1 && e->frame_next && n_steps > 0) { frame_at_end = true; e->frame_next = false; }
    const bool contact_on = e->prm.contact_flag >= 1 && !e->pairs.empty();
    if (e->prm.contact_dmax_clamp && (!e->halo.empty() || !e->g_node_map.empty()))
        return fail(e, HK_ERR_UNSUPPORTED, "contact_dmax_clamp: d_max is a global maximum; single-domain engines only");
    if (n_steps > 0) { int rc = ensure_erosion(e); if (rc) return rc; }
    const int64_t n_requested = n_steps;
#ifndef HK_EMU
    // replay: steps after the engine's first, without profiling events, clamp ping-pong, halos or a pending special case;
    // a frame's last step (it stores integ_triax_stress) and the remainder go through the plain loop below
    if (phase == 0 && !e->capturing && !e->graph_off && e->stream != nullptr && e->n_steps > 0 && !e->profiling &&
        e->halo.empty() && !e->comm && e->g_node_map.empty() && !e->prm.contact_dmax_clamp && !e->use_Q0 &&
        !e->contact_done && n_steps - (frame_at_end ? 1 : 0) >= 2) {
        if (e->graph_dirty || !e->step_graph) { int rc = step_graph_capture(e); if (rc) return rc; }
        if (e->step_graph) {
            const int64_t n_replay = n_steps - (frame_at_end ? 1 : 0);
            hk_launch_step_set(e->d_step, (long long)t_first, e->stream);
            for (int64_t i = 0; i < n_replay; ++i) CK(cudaGraphLaunch(e->step_graph, e->stream));
            e->n_launch += 1 + n_replay * e->graph_launches;
            e->n_steps += n_replay;
            if (e->any_ductile && e->dev_erosion) e->contact_host_stale = true;
            t_first += n_replay;
            n_steps -= n_replay;
        }
    }
#endif
    const HkDev& d = e->d;
    for (int64_t t = t_first; t < t_first + n_steps; ++t) {
        if (phase != 2 && contact_on && e->contact_done) {
            e->contact_done = false;             // done by hk_contact_enqueue (+ force exchange) for this step
        } else if (phase != 2 && contact_on) {
            int rc = contact_pass(e);
            if (rc) return rc;
        }
        unsigned long long* dmax_out = nullptr;       // clamp: this step's max |d_disp| becomes the next step's d_max
        if (contact_on && e->prm.contact_dmax_clamp && e->d_dmax) {
            dmax_out = e->d_dmax + (1 - e->clamp_cur);
            if (phase != 2) CK(hkp::dev_memset(dmax_out, 0, sizeof(unsigned long long), e->stream));
        }
        if (phase == 1) {
            prof_begin(e, 1);
            hk_launch_nodal(d, (double)t * e->prm.d_time, e->prm.d_time, e->dt2, e->dt2p, e->cp.lsb_exp,
                            contact_on ? 1 : 0, e->use_Q0, 1, nullptr, 0, dmax_out, e->stream);
            prof_end(e);
            e->n_launch += 1;
            continue;
        }
        if (!e->halo.empty()) {          // received partial forces + own -> halo_recv (ascending rank order)
            prof_begin(e, 3);
            int rc = halo_total(e);
            if (rc) return rc;
            prof_end(e);
        }
        prof_begin(e, 1);
        if (phase == 2)
            hk_launch_nodal(d, (double)t * e->prm.d_time, e->prm.d_time, e->dt2, e->dt2p, e->cp.lsb_exp,
                            contact_on ? 1 : 0, e->use_Q0, 2, e->d_halo_list, e->n_halo_nodes, dmax_out, e->stream);
        else
            hk_launch_nodal(d, (double)t * e->prm.d_time, e->prm.d_time, e->dt2, e->dt2p, e->cp.lsb_exp,
                            contact_on ? 1 : 0, e->use_Q0, 0, nullptr, 0, dmax_out, e->stream);
        prof_end(e);
        e->use_Q0 = 0;
        prof_begin(e, 2);
        CK(hk_launch_element(d, t, (frame_at_end && t == t_first + n_steps - 1) ? 1 : 0, e->stream));
        prof_end(e);
        e->n_launch += 2;
        if (e->any_ductile) {
            // marked elements: stress/strain zeroed; with contact their exposed faces join the surfaces ON THE DEVICE
            // (multi-GPU engines: the host driver replays the all-gathered ids through hk_apply_deleted instead)
            prof_begin(e, 5);
            long long nl = 0;
            if (e->dev_erosion && e->xerode)     // partitioned mesh: log + pack this rank's ids; every rank's are replayed below
                hk_launch_deletion_pass(d, nullptr, t, e->stream, &nl, e->xe_send, e->d_e_l2g, e->xe_cap);
            else
                hk_launch_deletion_pass(d, e->dev_erosion ? &e->er : nullptr, t, e->stream, &nl);
            e->n_launch += nl;
            prof_end(e);
            if (e->dev_erosion) e->contact_host_stale = true;
        }
        e->n_steps += 1;
        if (contact_on && e->prm.contact_dmax_clamp) e->clamp_cur = 1 - e->clamp_cur;
    }
    if (n_requested > 0 && phase != 1) {
        e->velo_current = contact_on;
        e->triax_current = frame_at_end;      // otherwise hk_download recomputes it from the current stress
    }
    return HK_OK;
}

#ifndef HK_EMU
// Multi-GPU steps with the engine's own communicator: per step
//     pack (main stream) -> ncclSend/ncclRecv with every neighbour (side stream) || nodal update of the non-interface
//     nodes (main stream) -> interface nodes, element kernel (main stream, after the exchange)
// all enqueued for n_steps steps without the host looking at anything in between.
// contact across ranks, all on the engine's stream (every stage needs the previous one): surface-node states of the
// nodes this rank owns -> ncclAllGather -> ghost copies -> contact pass on the local master triangles -> the 128-bit
// force accumulators as 43-bit limbs -> ncclAllReduce(int64, sum), exact -> accumulators identical on every rank
static int comm_contact_exchange(hk_engine* e) {
    if (!e->velo_current) { hk_launch_velo_from_rec(e->d, e->prm.d_time, e->stream); e->velo_current = true; }
    hk_launch_nodes_export(e->d, e->d_node_list[0], (long long)e->node_list[0].size(), e->cx_send, e->stream);
    NCK(g_nccl.AllGather(e->cx_send, e->cx_all, (size_t)e->cx_maxlen * 6, kNcclFloat64, e->comm, e->stream));
    hk_launch_nodes_import(e->d, e->d_node_list[1], e->d_import_src, (long long)e->node_list[1].size(), e->cx_all, e->stream);
    int rc = contact_pass(e);
    if (rc) return rc;
    e->contact_done = true;
    const long long ns = (long long)e->node_list[2].size();
    hk_launch_cacc_export_limbs(e->d, e->d_node_list[2], ns, e->cx_limbs, e->stream);
    NCK(g_nccl.AllReduce(e->cx_limbs, e->cx_limbs, (size_t)ns * 9, kNcclInt64, kNcclSum, e->comm, e->stream));
    hk_launch_cacc_import_limbs(e->d, e->d_node_list[2], ns, e->cx_limbs, e->stream);
    e->n_launch += 4;
    return 0;
}

// contact surfaces that erode across ranks (hk_comm_erosion): the step's deletions of every rank, as global ids, in one
// fixed-size all-gather; one thread replays them in ascending global order on every rank (add_surface_triangle,
// J2:767-804 + 2167-2245), so the pair lists of the next contact pass are in place without the host
static int comm_erosion_replay(hk_engine* e) {
    if (!e->dev_erosion || !e->xerode) return 0;
    prof_begin(e, 5);
    NCK(g_nccl.AllGather(e->xe_send, e->xe_all, (size_t)e->xe_cap + 1, kNcclInt64, e->comm, e->stream));
    hk_launch_erode_replay(e->d, e->er, e->xe_all, e->comm_world, e->xe_cap, e->stream);
    prof_end(e);
    e->n_launch += 1;
    e->contact_host_stale = true;
    return 0;
}

static int comm_steps(hk_engine* e, int64_t t_first, int64_t n_steps, bool frame_at_end) {
    if (e->frame_next && n_steps > 0) { frame_at_end = true; e->frame_next = false; }      // hk_mark_frame: LAST step
    const bool contact_on = e->prm.contact_flag >= 1 && !e->pairs.empty();
    for (int64_t t = t_first; t < t_first + n_steps; ++t) {
        int rc;
        if (contact_on && !e->contact_done) { if ((rc = comm_contact_exchange(e))) return rc; }
        if (e->halo.empty()) {                            // a rank with no interface node (does not occur with blocks)
            if (frame_at_end && t == t_first + n_steps - 1) e->frame_next = true;
            if ((rc = enqueue_steps(e, t, 1, false, 0))) return rc;
            if ((rc = comm_erosion_replay(e))) return rc;
            continue;
        }
        if ((rc = halo_pack_all(e))) return rc;
        CK(cudaEventRecord(e->ev_pack, e->stream));
        CK(cudaStreamWaitEvent(e->comm_stream, e->ev_pack, 0));
        prof_begin(e, 4, e->comm_stream, true);
        NCK(g_nccl.GroupStart());
        for (HaloNbr& h : e->halo) {
            NCK(g_nccl.Send(h.send, 3 * h.nodes.size(), kNcclFloat64, (int)h.rank, e->comm, e->comm_stream));
            NCK(g_nccl.Recv(h.recv, 3 * h.nodes.size(), kNcclFloat64, (int)h.rank, e->comm, e->comm_stream));
        }
        NCK(g_nccl.GroupEnd());
        prof_end(e, e->comm_stream, true);
        CK(cudaEventRecord(e->ev_comm, e->comm_stream));
        if ((rc = enqueue_steps(e, t, 1, false, 1))) return rc;       // overlaps the exchange
        CK(cudaStreamWaitEvent(e->stream, e->ev_comm, 0));
        if (frame_at_end && t == t_first + n_steps - 1) e->frame_next = true;
        if ((rc = enqueue_steps(e, t, 1, false, 2))) return rc;
        if ((rc = comm_erosion_replay(e))) return rc;
    }
    return 0;
}
#endif

static int step_enqueue_impl(hk_engine* e, int64_t t_first, int64_t n_steps, bool frame_at_end) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    if (n_steps < 0 || t_first < 0 || t_first + n_steps >= (1ll << 31)) return fail(e, HK_ERR_ARG, "bad step range");
#ifndef HK_EMU
    if (e->comm && (!e->halo.empty() || e->cx_maxlen > 0)) {   // the engine exchanges by itself: any number of steps
        if (e->prm.contact_flag >= 1 && !e->pairs.empty()) {
            if (e->cx_maxlen <= 0)
                return fail(e, HK_ERR_STATE, "contact across ranks: call hk_comm_contact after hk_set_node_list (or drive the "
                                             "exchange from the host: hk_nodes_* / hk_contact_* + hk_step_begin / hk_step_finish)");
            if (n_steps > 0) { int rc = ensure_erosion(e); if (rc) return rc; }
            if (n_steps > 1 && e->any_ductile && !e->g_node_map.empty() && !(e->dev_erosion && e->xerode))
                return fail(e, HK_ERR_UNSUPPORTED, "contact surfaces that erode across ranks: without hk_comm_erosion the host replays "
                                                   "the all-gathered deletions (hk_apply_deleted) after every step, so enqueue one step at a time");
        }
        int rc = comm_steps(e, t_first, n_steps, frame_at_end);
        if (rc) return rc;
        CK(hkp::last_error());
        return HK_OK;
    }
#endif
    if (!e->halo.empty() && n_steps > 1)
        return fail(e, HK_ERR_ARG, "with halos and no communicator (hk_comm_init), exchange and step one step at a time");

    int rc = enqueue_steps(e, t_first, n_steps, frame_at_end, 0);
    if (rc) return rc;
    CK(hkp::last_error());
    return HK_OK;
}

// asynchronous form: no output frame is implied, so integ_triax_stress is not stored (hk_download derives it from
// the stress when asked)
int HKAPI(step_enqueue)(hk_engine* e, int64_t t_first, int64_t n_steps) {
    return step_enqueue_impl(e, t_first, n_steps, false);
}

// split step for the multi-GPU driver: hk_step_begin(t) runs what does not depend on the halo (so it overlaps the
// NCCL exchange), hk_step_finish(t) the rest.  hk_step_begin + hk_step_finish == hk_step_enqueue(t, 1).
// The asynchronous forms imply no output frame.  hk_mark_frame says the last step of the NEXT hk_step_enqueue /
// hk_step_finish call is followed by one: that step stores integ_triax_stress as computed inside it (J2:677),
// i.e. before the fracture pass zeroes the stress of the elements it deletes (what a frame of the reference shows).
int HKAPI(mark_frame)(hk_engine* e) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    e->frame_next = true;
    return HK_OK;
}

int HKAPI(step_begin)(hk_engine* e, int64_t t) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    if (e->begun_t >= 0) return fail(e, HK_ERR_STATE, "hk_step_begin called twice without hk_step_finish");
    int rc = enqueue_steps(e, t, 1, false, 1);
    if (rc) return rc;
    e->begun_t = t;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(step_finish)(hk_engine* e, int64_t t) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    if (e->begun_t != t) return fail(e, HK_ERR_STATE, "hk_step_finish without matching hk_step_begin");
    e->begun_t = -1;
    int rc = enqueue_steps(e, t, 1, false, 2);
    if (rc) return rc;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(sync)(hk_engine* e, int64_t* n_deleted_out) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    const size_t before = e->deleted_reported;
    int rc = fetch_deleted(e, nullptr);
    if (rc) return rc;
    CK(hkp::sync(e->stream));
    CK(hkp::last_error());
    e->deleted_reported = e->deleted_all.size();
    if (n_deleted_out) *n_deleted_out = (int64_t)(e->deleted_all.size() - before);
    return HK_OK;
}

// synchronous form: the caller typically writes a frame after it, so the last step stores integ_triax_stress
int HKAPI(step)(hk_engine* e, int64_t t_first, int64_t n_steps, int64_t* n_deleted_out) {
    int rc = step_enqueue_impl(e, t_first, n_steps, true);
    if (rc) return rc;
    return HKAPI(sync)(e, n_deleted_out);
}

// staging buffer for layout transposes
static int ensure_staging(hk_engine* e, size_t doubles) {
    if (e->staging_doubles >= doubles) return 0;
    dfree(e, e->staging);
    e->staging = nullptr;
    int rc = dalloc(e, &e->staging, doubles);
    if (rc) return rc;
    e->staging_doubles = doubles;
    return 0;
}
static const long long CHUNK_E = 1 << 18;   // elements per transposed chunk (<= 100 MB per staging block)

static int ensure_stage2(hk_engine* e, size_t doubles) {
#ifndef HK_EMU
    if (!e->copy_stream) {
        CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            CK(cudaEventCreateWithFlags(&e->ev_ready[b], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e->ev_free[b], cudaEventDisableTiming));
        }
    }
#endif
    if (e->stage2_doubles >= doubles) return 0;
    for (int b = 0; b < 2; ++b) {
        dfree(e, e->stage2[b]);
        e->stage2[b] = nullptr;
        int rc = dalloc(e, &e->stage2[b], doubles);
        if (rc) return rc;
    }
    e->stage2_doubles = doubles;
    return 0;
}

// Gauss-point fields between the reference's (ncomp, nip) arrays and the tile-blocked device rows, chunk by chunk
// through two staging blocks: the PCIe copy of chunk i+1 (copy stream) overlaps the transpose of chunk i (engine
// stream); events order them, the host never waits inside the loop.  The caller synchronises both streams once.
static int download_ip(hk_engine* e, int row0, double* host, int ncomp) {   // row0 < 0: triax
    if (!host) return 0;
    const long long nE = e->nElement;
    int rc = ensure_stage2(e, (size_t)std::min<long long>(CHUNK_E, nE) * 8 * 6);
    if (rc) return rc;
    long long i = 0;
    for (long long e0 = 0; e0 < nE; e0 += CHUNK_E, ++i) {
        const int b = (int)(i & 1);
        const long long ne = std::min<long long>(CHUNK_E, nE - e0);
#ifndef HK_EMU
        if (i >= 2) CK(cudaStreamWaitEvent(e->stream, e->ev_free[b], 0));        // its previous copy has left the block
#endif
        if (row0 < 0) hk_launch_triax_to_aos(e->d, e->stage2[b], e0, ne, e->stream);
        else hk_launch_ip_to_aos(e->d, e->stage2[b], row0, ncomp, e0, ne, e->stream);
#ifndef HK_EMU
        CK(cudaEventRecord(e->ev_ready[b], e->stream));
        CK(cudaStreamWaitEvent(e->copy_stream, e->ev_ready[b], 0));
        CK(hkp::d2h_async(host + e0 * 8 * ncomp, e->stage2[b], (size_t)ne * 8 * ncomp * sizeof(double), e->copy_stream));
        CK(cudaEventRecord(e->ev_free[b], e->copy_stream));
#else
        CK(hkp::d2h(host + e0 * 8 * ncomp, e->stage2[b], (size_t)ne * 8 * ncomp * sizeof(double), e->stream));
#endif
    }
#ifndef HK_EMU
    CK(cudaStreamSynchronize(e->copy_stream));       // the next field reuses the blocks from chunk 0
#endif
    return 0;
}
static int upload_ip(hk_engine* e, int row0, const double* host, int ncomp) {
    if (!host) return 0;
    const long long nE = e->nElement;
    int rc = ensure_stage2(e, (size_t)std::min<long long>(CHUNK_E, nE) * 8 * 6);
    if (rc) return rc;
    long long i = 0;
    for (long long e0 = 0; e0 < nE; e0 += CHUNK_E, ++i) {
        const int b = (int)(i & 1);
        const long long ne = std::min<long long>(CHUNK_E, nE - e0);
#ifndef HK_EMU
        if (i >= 2) CK(cudaStreamWaitEvent(e->copy_stream, e->ev_free[b], 0));   // its previous transpose has read the block
        CK(hkp::h2d_async(e->stage2[b], host + e0 * 8 * ncomp, (size_t)ne * 8 * ncomp * sizeof(double), e->copy_stream));
        CK(cudaEventRecord(e->ev_ready[b], e->copy_stream));
        CK(cudaStreamWaitEvent(e->stream, e->ev_ready[b], 0));
#else
        CK(hkp::h2d(e->stage2[b], host + e0 * 8 * ncomp, (size_t)ne * 8 * ncomp * sizeof(double), e->stream));
#endif
        hk_launch_ip_to_dev(e->stage2[b], e->d, row0, ncomp, e0, ne, e->stream);
#ifndef HK_EMU
        CK(cudaEventRecord(e->ev_free[b], e->stream));
#endif
    }
#ifndef HK_EMU
    CK(cudaStreamSynchronize(e->stream));            // the next field reuses the blocks from chunk 0
#endif
    return 0;
}

int HKAPI(download)(hk_engine* e, double* disp, double* velo, double* integ_stress, double* integ_strain,
                    double* integ_eq_plastic_strain, double* integ_triax_stress, int64_t* element_flag) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    const HkDev& d = e->d;
    const size_t fnb = sizeof(double) * 3 * e->nNode;
    int rc;
    if (disp) CK(hkp::d2h_async(disp, d.u, fnb, e->stream));          // one synchronisation at the end of the call
    if (velo) {
        if (!e->velo_current) { hk_launch_velo_from_rec(d, e->prm.d_time, e->stream); e->velo_current = true; }
        CK(hkp::d2h_async(velo, d.velo, fnb, e->stream));
    }
    if ((rc = download_ip(e, 0, integ_stress, 6))) return rc;
    if ((rc = download_ip(e, 6, integ_strain, 6))) return rc;
    if ((rc = download_ip(e, 12, integ_eq_plastic_strain, 1))) return rc;
    if (integ_triax_stress) {
        if (!e->triax_current) { hk_launch_triax(d, e->stream); e->triax_current = true; }
        if ((rc = download_ip(e, -1, integ_triax_stress, 1))) return rc;
    }
    if (element_flag) {
        std::vector<unsigned char> fl(e->nElement);
        CK(hkp::d2h(fl.data(), d.flag, fl.size(), e->stream));
        for (int64_t i = 0; i < e->nElement; ++i) element_flag[i] = fl[i] == 1 ? 1 : 0;
    }
    CK(hkp::sync(e->stream));
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(node_output)(hk_engine* e, double* node_stress, double* node_strain, double* node_eq_plastic_strain,
                       double* node_mises_stress, double* node_triax_stress, double* inc_num, int32_t raw) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    const HkDev& d = e->d;
    if (!e->triax_current) { hk_launch_triax(d, e->stream); e->triax_current = true; }
    int rc = 0;
    if (!e->no_emean) rc = dalloc(e, &e->no_emean, (size_t)14 * d.nEp);
    if (!rc && !e->no_out) rc = dalloc(e, &e->no_out, (size_t)16 * d.nNode);
    if (rc) return rc;
    double* emean = e->no_emean;
    double* out = e->no_out;
    hk_launch_element_means(d, emean, e->stream);
    hk_launch_node_means(d, emean, out, raw ? 1 : 0, e->stream);
    e->n_launch += 2;
    const size_t nb = sizeof(double) * (size_t)d.nNode;
    int err = 0;
    if (node_stress) err |= hkp::d2h_async(node_stress, out, 6 * nb, e->stream);
    if (node_strain) err |= hkp::d2h_async(node_strain, out + 6 * d.nNode, 6 * nb, e->stream);
    if (node_eq_plastic_strain) err |= hkp::d2h_async(node_eq_plastic_strain, out + 12 * d.nNode, nb, e->stream);
    if (node_mises_stress && !raw) err |= hkp::d2h_async(node_mises_stress, out + 13 * d.nNode, nb, e->stream);
    if (node_triax_stress) err |= hkp::d2h_async(node_triax_stress, out + 14 * d.nNode, nb, e->stream);
    if (inc_num) err |= hkp::d2h_async(inc_num, out + 15 * d.nNode, nb, e->stream);
    err |= hkp::sync(e->stream);
    if (err) return fail(e, HK_ERR_CUDA, "hk_node_output: copy failed");
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(download_ex)(hk_engine* e, double* disp_pre, double* Q, double* external_force, double* position,
                       double* integ_yield_stress, double* elementVolume) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    const HkDev& d = e->d;
    const int64_t nN = e->nNode;
    const size_t fnb = sizeof(double) * 3 * nN;
    int rc;
    if (disp_pre) CK(hkp::d2h(disp_pre, d.u_pre, fnb, e->stream));
    if (Q) {
        if (e->use_Q0) {
            CK(hkp::d2h(Q, d.Q0, fnb, e->stream));
        } else {
            if ((rc = ensure_staging(e, (size_t)3 * nN))) return rc;
            hk_launch_gather_Q(d, e->staging, e->stream);
            CK(hkp::d2h(Q, e->staging, fnb, e->stream));
        }
    }
    if (external_force) {
        if ((rc = ensure_staging(e, (size_t)3 * nN))) return rc;
        double* out = e->staging;
        const HkDev dd = d;
        const int lsb = e->cp.lsb_exp;
        const int on = (e->prm.contact_flag >= 1 && !e->pairs.empty()) ? 1 : 0;
        hk_launch_external_force(dd, out, lsb, on, e->stream);
        CK(hkp::d2h(external_force, out, fnb, e->stream));
    }
    if (position) {
        std::vector<double> rec(6 * nN);
        CK(hkp::d2h(rec.data(), d.rec, rec.size() * sizeof(double), e->stream));
        for (int64_t n = 0; n < nN; ++n)
            for (int c = 0; c < 3; ++c) position[3 * n + c] = rec[6 * n + c];
    }
    if ((rc = download_ip(e, 13, integ_yield_stress, 1))) return rc;
    if (elementVolume) {
        if ((rc = ensure_staging(e, (size_t)e->nElement))) return rc;
        hk_launch_element_volume(d, e->staging, e->stream);
        CK(hkp::d2h(elementVolume, e->staging, sizeof(double) * e->nElement, e->stream));
    }
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(upload_state)(hk_engine* e, const double* disp, const double* disp_pre, const double* velo, const double* Q,
                        const double* integ_stress, const double* integ_strain, const double* integ_eq_plastic_strain,
                        const double* integ_yield_stress, const int64_t* element_flag) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    HkDev& d = e->d;
    const int64_t nN = e->nNode;
    const size_t fnb = sizeof(double) * 3 * nN;
    int rc;
    if (disp) {
        CK(hkp::h2d_async(d.u, disp, fnb, e->stream));                // synchronised once at the end of the call
        const HkDev dd = d;
        hk_parallel_for(nN * 3, e->stream, HK_LAMBDA(long long i) {      // position = coordmat + disp, J2:650-652
            const long long n = i / 3;
            const int c = (int)(i - 3 * n);
            dd.rec[6 * n + c] = dd.X[i] + dd.u[i];
        });
    }
    if (disp_pre) CK(hkp::h2d_async(d.u_pre, disp_pre, fnb, e->stream));
    if (velo) { CK(hkp::h2d_async(d.velo, velo, fnb, e->stream)); e->velo_current = true; }
    if (Q) { CK(hkp::h2d_async(d.Q0, Q, fnb, e->stream)); e->use_Q0 = 1; }
    if ((rc = upload_ip(e, 0, integ_stress, 6))) return rc;
    if ((rc = upload_ip(e, 6, integ_strain, 6))) return rc;
    if ((rc = upload_ip(e, 12, integ_eq_plastic_strain, 1))) return rc;
    if ((rc = upload_ip(e, 13, integ_yield_stress, 1))) return rc;
    if (integ_stress) e->triax_current = false;
    if (element_flag) {
        std::vector<unsigned char> fl(e->nElement);
        for (int64_t i = 0; i < e->nElement; ++i) fl[i] = element_flag[i] == 1 ? 1 : 2;
        CK(hkp::h2d(d.flag, fl.data(), fl.size(), e->stream));
        CK(hkp::dev_memset(d.Qe, 0, sizeof(double) * 24 * d.nEp, e->stream));   // supply Q with the flags
    }
    CK(hkp::sync(e->stream));
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(deleted_ids)(hk_engine* e, int64_t* ids, int64_t cap, int64_t* n_out) {
    if (!e) return HK_ERR_ARG;
    const int64_t n = (int64_t)e->deleted_all.size();
    if (n_out) *n_out = n;
    if (ids) for (int64_t i = 0; i < std::min(n, cap); ++i) ids[i] = e->deleted_all[i];
    return HK_OK;
}

int HKAPI(deleted_steps)(hk_engine* e, int64_t* steps, int64_t cap, int64_t* n_out) {
    if (!e) return HK_ERR_ARG;
    const int64_t n = (int64_t)e->deleted_step.size();
    if (n_out) *n_out = n;
    if (steps) for (int64_t i = 0; i < std::min(n, cap); ++i) steps[i] = e->deleted_step[i];
    return HK_OK;
}

int HKAPI(contact_pair_info)(hk_engine* e, int64_t c, int64_t* nn_i, int64_t* nn_j, int64_t* nTri, int64_t* c_nodes_i,
                             int64_t* c_nodes_j, int64_t* c_triangles, int64_t* c_triangles_eleid) {
    if (!e || c < 0 || c >= (int64_t)e->pairs.size()) return fail(e, HK_ERR_ARG, "bad contact pair index");
    { int rc = contact_refresh_host(e); if (rc) return rc; }
    const PairH& p = e->pairs[c];
    const int64_t nt = (int64_t)p.t0.size();
    if (nn_i) *nn_i = (int64_t)p.nodes_i.size();
    if (nn_j) *nn_j = (int64_t)p.nodes_j.size();
    if (nTri) *nTri = nt;
    if (c_nodes_i) for (size_t k = 0; k < p.nodes_i.size(); ++k) c_nodes_i[k] = p.nodes_i[k] + 1;
    if (c_nodes_j) for (size_t k = 0; k < p.nodes_j.size(); ++k) c_nodes_j[k] = p.nodes_j[k] + 1;
    if (c_triangles)
        for (int64_t k = 0; k < nt; ++k) {
            c_triangles[k] = p.t0[k] + 1;
            c_triangles[k + nt] = p.t1[k] + 1;
            c_triangles[k + 2 * nt] = p.t2[k] + 1;
        }
    if (c_triangles_eleid) for (int64_t k = 0; k < nt; ++k) c_triangles_eleid[k] = p.tele[k] + 1;
    return HK_OK;
}

int HKAPI(counters)(hk_engine* e, int64_t out[8]) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    unsigned long long c[8];
    CK(hkp::d2h(c, e->d.counters, sizeof(c), e->stream));
    out[0] = (int64_t)c[0];
    out[1] = (int64_t)c[1];
    out[2] = (int64_t)c[2];
    out[3] = e->n_launch;
    out[4] = e->n_steps;
    out[5] = (int64_t)c[3];
    out[6] = 0;
    out[7] = 0;
    return HK_OK;
}

int HKAPI(state_summary)(hk_engine* e, double out[8]) {
    if (!e || !e->finalized || !out) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    if (!e->d_summary) { int rc = dalloc(e, &e->d_summary, (size_t)4); if (rc) return rc; }
    const unsigned long long init[4] = {0ull, ~0ull, 0ull, 0ull};
    unsigned long long got[4];
    CK(hkp::h2d(e->d_summary, init, sizeof(init), e->stream));
    hk_launch_state_summary(e->d, e->d_summary, e->stream);
    e->n_launch += 1;
    CK(hkp::d2h(got, e->d_summary, sizeof(got), e->stream));
    CK(hkp::last_error());
    for (int i = 0; i < 8; ++i) out[i] = 0.0;
    out[0] = (double)got[0];
    out[1] = got[0] ? hk_decode_double(got[1]) : 0.0;
    out[2] = got[0] ? hk_decode_double(got[2]) : 0.0;
    out[3] = (double)got[3];
    return HK_OK;
}

int HKAPI(profile)(hk_engine* e, int32_t enable) {
    if (!e) return HK_ERR_ARG;
    prof_collect(e);
    for (int i = 0; i < 8; ++i) { e->prof_ms[i] = 0; e->prof_n[i] = 0; }
    e->profiling = enable != 0;
    return HK_OK;
}

int HKAPI(profile_read)(hk_engine* e, double ms[4], int64_t launches[4]) {
    if (!e) return HK_ERR_ARG;
    prof_collect(e);
    for (int i = 0; i < 4; ++i) { ms[i] = e->prof_ms[i]; launches[i] = e->prof_n[i]; }
    return HK_OK;
}

int HKAPI(profile_read_ex)(hk_engine* e, double ms[8], int64_t launches[8]) {
    if (!e) return HK_ERR_ARG;
    prof_collect(e);
    for (int i = 0; i < 8; ++i) { ms[i] = e->prof_ms[i]; launches[i] = e->prof_n[i]; }
    return HK_OK;
}

int HKAPI(set_halo)(hk_engine* e, int64_t n_neighbors, const int64_t* nbr_ptr, const int64_t* nodes) {
    if (!e) return HK_ERR_ARG;
    if (e->finalized) return fail(e, HK_ERR_STATE, "hk_set_halo must precede hk_finalize");
    e->halo.clear();
    for (int64_t i = 0; i < n_neighbors; ++i) {
        HaloNbr h;
        for (int64_t k = nbr_ptr[i]; k < nbr_ptr[i + 1]; ++k) h.nodes.push_back((int)(nodes[k] - 1));
        e->halo.push_back(std::move(h));
    }
    return HK_OK;
}

int HKAPI(comm_contact)(hk_engine* e, int64_t maxlen, const int64_t* src_index) {
#ifndef HK_EMU
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    if (!e->comm) return fail(e, HK_ERR_STATE, "hk_comm_init first");
    if (maxlen < 1 || (int64_t)e->node_list[0].size() > maxlen) return fail(e, HK_ERR_ARG, "maxlen smaller than this rank's export list");
    const size_t n_ghost = e->node_list[1].size();
    if (n_ghost && !src_index) return fail(e, HK_ERR_ARG, "src_index missing");
    for (size_t i = 0; i < n_ghost; ++i)
        if (src_index[i] < 0 || src_index[i] >= maxlen * e->comm_world) return fail(e, HK_ERR_ARG, "src_index out of range");
    CK(hkp::sync(e->stream));
    dfree(e, e->cx_send); dfree(e, e->cx_all); dfree(e, e->cx_limbs); dfree(e, e->d_import_src);
    e->cx_send = nullptr; e->cx_all = nullptr; e->cx_limbs = nullptr; e->d_import_src = nullptr;
    int rc;
    if ((rc = dalloc(e, &e->cx_send, (size_t)maxlen * 6))) return rc;
    if ((rc = dalloc(e, &e->cx_all, (size_t)maxlen * 6 * e->comm_world))) return rc;
    if ((rc = dalloc(e, &e->cx_limbs, std::max<size_t>(1, e->node_list[2].size() * 9)))) return rc;
    CK(hkp::dev_memset(e->cx_send, 0, (size_t)maxlen * 6 * sizeof(double), e->stream));
    std::vector<long long> src(src_index, src_index + n_ghost);
    if ((rc = dalloc(e, &e->d_import_src, std::max<size_t>(1, n_ghost)))) return rc;
    if ((rc = upload(e, e->d_import_src, src))) return rc;
    e->cx_maxlen = maxlen;
    return HK_OK;
#else
    (void)maxlen; (void)src_index;
    return fail(e, HK_ERR_UNSUPPORTED, "host-compiled debugging build has no NCCL");
#endif
}

int HKAPI(halo_bind)(hk_engine* e, int64_t neighbor, void* send_dev, void* recv_dev) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    if (neighbor < 0 || neighbor >= (int64_t)e->halo.size()) return fail(e, HK_ERR_ARG, "bad neighbour index");
    e->halo[neighbor].send = (double*)send_dev;
    e->halo[neighbor].recv = (double*)recv_dev;
    return HK_OK;
}

int HKAPI(halo_pack)(hk_engine* e) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    int rc = halo_pack_all(e);
    if (rc) return rc;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(set_halo_ranks)(hk_engine* e, int64_t my_rank, int64_t n_neighbors, const int64_t* ranks) {
    if (!e) return HK_ERR_ARG;
    if (n_neighbors != (int64_t)e->halo.size()) return fail(e, HK_ERR_ARG, "hk_set_halo_ranks: one rank per hk_set_halo neighbour");
    if (my_rank < 0) return fail(e, HK_ERR_ARG, "bad rank");
    for (int64_t i = 0; i < n_neighbors; ++i) {
        if (!ranks || ranks[i] < 0 || ranks[i] == my_rank) return fail(e, HK_ERR_ARG, "bad neighbour rank");
        e->halo[i].rank = ranks[i];
    }
    e->my_rank = my_rank;
    return HK_OK;
}

int HKAPI(comm_unique_id)(void* id128) {
#ifndef HK_EMU
    if (!id128) return HK_ERR_ARG;
    if (!nccl_load()) return fail(nullptr, HK_ERR_UNSUPPORTED, g_nccl.err);
    HkNcclId id;
    const int rn = g_nccl.GetUniqueId(&id);
    if (rn) return fail(nullptr, HK_ERR_CUDA, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rn));
    std::memcpy(id128, &id, sizeof(id));
    return HK_OK;
#else
    (void)id128;
    return fail(nullptr, HK_ERR_UNSUPPORTED, "host-compiled debugging build has no NCCL");
#endif
}

int HKAPI(comm_init)(hk_engine* e, const void* id128, int32_t rank, int32_t world) {
#ifndef HK_EMU
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(e, HK_ERR_ARG, "bad communicator arguments");
    if (e->comm) return fail(e, HK_ERR_STATE, "communicator already created");
    if (!e->halo.empty() && (e->my_rank != rank)) return fail(e, HK_ERR_STATE, "hk_set_halo_ranks must give this rank before hk_comm_init");
    for (const HaloNbr& h : e->halo)
        if (h.rank < 0 || h.rank >= world) return fail(e, HK_ERR_STATE, "hk_set_halo_ranks: neighbour rank outside the communicator");
    if (!nccl_load()) return fail(e, HK_ERR_UNSUPPORTED, g_nccl.err);
    CK(cudaSetDevice(e->prm.device));
    HkNcclId id;
    std::memcpy(&id, id128, sizeof(id));
    NCK(g_nccl.CommInitRank(&e->comm, world, id, rank));
    e->comm_world = world;
    CK(cudaStreamCreateWithFlags(&e->comm_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_pack, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_comm, cudaEventDisableTiming));
    for (HaloNbr& h : e->halo) {                      // the engine's own exchange buffers
        if (h.send && h.recv) continue;
        int rc;
        if ((rc = dalloc(e, &h.send, 3 * h.nodes.size()))) return rc;
        if ((rc = dalloc(e, &h.recv, 3 * h.nodes.size()))) return rc;
        h.own_buffers = true;
    }
    return HK_OK;
#else
    (void)id128; (void)rank; (void)world;
    return fail(e, HK_ERR_UNSUPPORTED, "host-compiled debugging build has no NCCL");
#endif
}

int HKAPI(set_node_list)(hk_engine* e, int32_t which, int64_t n, const int64_t* nodes) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    if (which < 0 || which > 4) return fail(e, HK_ERR_ARG, "bad list id");
    { int rc = contact_refresh_host(e); if (rc) return rc; }
    std::vector<int>& L = e->node_list[which];
    L.resize(n);
    for (int64_t i = 0; i < n; ++i) {
        if (nodes[i] < 1 || nodes[i] > e->nNode) return fail(e, HK_ERR_ARG, "node id out of range");
        L[i] = (int)(nodes[i] - 1);
        if (which == 2 && !e->xerode && (e->spec_idx_h[L[i]] < 0 || e->spec_h[e->spec_idx_h[L[i]]].contact_slot < 0))
            return fail(e, HK_ERR_ARG, "surface list holds a node that is in no contact pair");
    }
    dfree(e, e->d_node_list[which]);
    e->d_node_list[which] = nullptr;
    int rc = dalloc(e, &e->d_node_list[which], L.size());
    if (rc) return rc;
    return upload(e, e->d_node_list[which], L);
}

int HKAPI(nodes_export)(hk_engine* e, void* out_dev) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    if (!e->velo_current) { hk_launch_velo_from_rec(e->d, e->prm.d_time, e->stream); e->velo_current = true; }
    hk_launch_nodes_export(e->d, e->d_node_list[0], (long long)e->node_list[0].size(), (double*)out_dev, e->stream);
    e->n_launch += 1;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(nodes_import)(hk_engine* e, const void* in_dev, const int64_t* src_index) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    const size_t n = e->node_list[1].size();
    if (src_index) {               // (re)register where each ghost's record sits in the gathered buffer
        std::vector<long long> src(src_index, src_index + n);
        dfree(e, e->d_import_src);
        e->d_import_src = nullptr;
        int rc = dalloc(e, &e->d_import_src, n);
        if (rc) return rc;
        if ((rc = upload(e, e->d_import_src, src))) return rc;
    }
    if (n && !e->d_import_src) return fail(e, HK_ERR_STATE, "hk_nodes_import: source index never given");
    hk_launch_nodes_import(e->d, e->d_node_list[1], e->d_import_src, (long long)n, (const double*)in_dev, e->stream);
    e->n_launch += 1;
    CK(hkp::last_error());
    return HK_OK;
}

// ghost-element partitions (bit-identical results for any number of ranks): displacement state of listed nodes
int HKAPI(state_export)(hk_engine* e, void* out_dev) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    hk_launch_state_export(e->d, e->d_node_list[3], (long long)e->node_list[3].size(), (double*)out_dev, e->stream);
    e->n_launch += 1;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(state_import)(hk_engine* e, const void* in_dev) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    hk_launch_state_import(e->d, e->d_node_list[4], (long long)e->node_list[4].size(), (const double*)in_dev,
                           e->prm.d_time, e->stream);
    e->n_launch += 1;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(comm_erosion)(hk_engine* e, int32_t max_deleted_per_step) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    if (e->g_node_map.empty()) return fail(e, HK_ERR_STATE, "hk_comm_erosion: call hk_set_global_maps first");
    if (e->erosion_checked) return fail(e, HK_ERR_STATE, "hk_comm_erosion: call before the first step");
    if (max_deleted_per_step < 1 || max_deleted_per_step > (1 << 24)) return fail(e, HK_ERR_ARG, "bad capacity");
    e->xerode = true;
    e->xe_cap = max_deleted_per_step;
    return HK_OK;
}

int HKAPI(set_global_maps)(hk_engine* e, int64_t n_global_nodes, const int64_t* node_map, int64_t n_global_elements,
                           const int64_t* elem_map, const int64_t* element_instance) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    if (!node_map || !elem_map || !element_instance) return fail(e, HK_ERR_ARG, "null argument");
    e->g_node_map.resize(n_global_nodes);
    for (int64_t i = 0; i < n_global_nodes; ++i) {
        if (node_map[i] > e->nNode) return fail(e, HK_ERR_ARG, "node_map entry out of range");
        e->g_node_map[i] = (int)(node_map[i] - 1);              // 0 (absent) -> -1
    }
    e->g_elem_map.resize(n_global_elements);
    e->g_einst.assign(element_instance, element_instance + n_global_elements);
    for (int64_t i = 0; i < n_global_elements; ++i) {
        if (elem_map[i] > e->nElement) return fail(e, HK_ERR_ARG, "elem_map entry out of range");
        e->g_elem_map[i] = (int)(elem_map[i] - 1);
    }
    return HK_OK;
}

int HKAPI(apply_deleted)(hk_engine* e, int64_t n, const int64_t* global_ids) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    const bool global = !e->g_node_map.empty();
    const int64_t limit = global ? (int64_t)e->g_elem_map.size() : e->nElement;
    std::vector<int64_t> ids(global_ids, global_ids + n);
    for (int64_t g : ids)
        if (g < 1 || g > limit) return fail(e, HK_ERR_ARG, "element id out of range");
    int rc = (global && !e->xerode) ? 0 : ensure_erosion(e);   // restart hook: the replay and later steps share one set of lists
    if (rc) return rc;
    if (global && e->dev_erosion && e->xerode) {      // host-driven exchange, device-side lists: replay the ids on the device
        if (e->comm) return fail(e, HK_ERR_STATE, "hk_comm_erosion with a communicator: the engine replays deletions itself");
        std::vector<long long> buf((size_t)e->xe_cap + 1);
        for (int64_t done = 0; done < n;) {
            const int64_t m = std::min<int64_t>(n - done, e->xe_cap);
            buf[0] = m;
            for (int64_t i = 0; i < m; ++i) buf[1 + i] = ids[done + i] - 1;
            CK(hkp::h2d(e->xe_all, buf.data(), (size_t)(m + 1) * sizeof(long long), e->stream));
            hk_launch_erode_replay(e->d, e->er, e->xe_all, 1, e->xe_cap, e->stream);
            CK(hkp::sync(e->stream));
            done += m;
        }
        e->n_launch += 1;
        e->contact_host_stale = true;
        CK(hkp::last_error());
        return HK_OK;
    }
    if ((rc = contact_refresh_host(e))) return rc;
    rc = update_surfaces(e, ids);
    if (rc) return rc;
    if (!global) {                       // restart of a single-domain run: the replayed ids are part of the history
        e->deleted_all.insert(e->deleted_all.end(), ids.begin(), ids.end());
        e->deleted_step.insert(e->deleted_step.end(), ids.size(), 0);
        e->deleted_reported = e->deleted_all.size();
    }
    CK(hkp::sync(e->stream));
    return HK_OK;
}

int HKAPI(contact_enqueue)(hk_engine* e) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    const bool contact_on = e->prm.contact_flag >= 1 && !e->pairs.empty();
    if (!contact_on) return HK_OK;
    { int rc = ensure_erosion(e); if (rc) return rc; }
    { int rc = contact_pass(e); if (rc) return rc; }
    e->contact_done = true;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(contact_export)(hk_engine* e, void* out_dev) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    hk_launch_cacc_export(e->d, e->d_node_list[2], (long long)e->node_list[2].size(), (unsigned long long*)out_dev, e->stream);
    e->n_launch += 1;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(contact_import)(hk_engine* e, const void* in_dev, int64_t n_ranks) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    hk_launch_cacc_import(e->d, e->d_node_list[2], (long long)e->node_list[2].size(), (const unsigned long long*)in_dev,
                          (long long)n_ranks, e->stream);
    e->n_launch += 1;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(contact_export_limbs)(hk_engine* e, void* out_dev) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    hk_launch_cacc_export_limbs(e->d, e->d_node_list[2], (long long)e->node_list[2].size(), (long long*)out_dev, e->stream);
    e->n_launch += 1;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(contact_import_limbs)(hk_engine* e, const void* in_dev) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    HK_DEVICE(e);
    hk_launch_cacc_import_limbs(e->d, e->d_node_list[2], (long long)e->node_list[2].size(), (const long long*)in_dev, e->stream);
    e->n_launch += 1;
    CK(hkp::last_error());
    return HK_OK;
}

int HKAPI(set_stream)(hk_engine* e, void* cuda_stream) {
    if (!e) return HK_ERR_ARG;
#ifndef HK_EMU
    hkp::sync(e->stream);
    if (e->own_stream) { cudaStreamDestroy(e->stream); e->own_stream = false; }
    e->stream = (cudaStream_t)cuda_stream;      // NULL is a valid handle: the CUDA legacy default stream
    e->graph_dirty = true;                      // (the legacy stream cannot be captured: steps on it are plain launches)
#else
    (void)cuda_stream;
#endif
    return HK_OK;
}

}  // extern "C"
