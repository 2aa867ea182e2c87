// hk_common.h — device-side data layout shared by the engine's translation units.
//
// Layout in HBM:
//   nodal vectors        AoS, 3 doubles per node, identical to the reference's `fn` vectors
//                        (disp, disp_pre, coordmat): perfectly coalesced for the streaming update.
//   node record `rec`    {x,y,z, dux,duy,duz} per node (48 B, 16-B aligned): `position` and `d_disp`
//                        of J2:624-652, written by the nodal kernel, gathered by the element kernel.
//   ip state             tile-blocked SoA [tile][gauss point][row 14][TL]: the state of one Gauss point of one tile
//                        of TL consecutive elements is a single contiguous burst (39 KB at TL = 352) that the
//                        TMA moves with one bulk copy; within it consecutive elements are consecutive doubles.
//   Qe                   SoA [24][nEp]: element nodal forces (J2:438), written by the element kernel and
//                        gathered per node in ascending element order (= the reference's serial scatter
//                        order, J2:669-675) by the nodal kernel of the next step.  Q itself is never
//                        materialised.
#pragma once
#include "hk_platform.h"

#define HK_MAX_TABLE 32          // rows of a *Plastic / ductile table

struct HkMaterialDev {
    double young, poisson, G;
    double D11, D12, D44;        // entries of Dmat, built exactly as J2:143-159
    int npp, nd;
    double plastic_s[HK_MAX_TABLE];   // yield stress column
    double plastic_e[HK_MAX_TABLE];   // eq. plastic strain column
    double Hd[HK_MAX_TABLE];
    double duct_e[HK_MAX_TABLE];      // fracture strain column
    double duct_t[HK_MAX_TABLE];      // triaxiality column
};

// special-node table entry: everything rare that the nodal kernel must look at
struct HkSpecialNode {
    int bc_entry[3];             // per dof: index into bc_value/bc_amp, or -1
    int contact_slot;            // index into the contact force accumulators, or -1
    int halo_slot;               // index into the received halo partial forces, or -1
    int pad;
};

struct HkAmpTable {              // one per BC that has an amplitude (J2:586-600)
    int n;                       // number of points
    int offset;                  // into amp_time / amp_value
};

struct HkDev {
    long long nNode, nElement, nEp;   // nEp: padded element count (row stride of the SoA arrays)
    int ell_width;                    // max elements per node
    int n_mat;
    // nodes
    double* X;            // [nNode*3] coordmat
    double* u;            // disp
    double* u_pre;        // disp_pre
    double* velo;         // velo (kept current only when contact is on; see hk_engine.cu)
    double* rec;          // [nNode*6] {position, d_disp}
    double* mass;         // [nNode]
    int* ell;             // [ell_width][nNode] entries e*8+a (ascending e), -1 = empty
    int* spec_idx;        // [nNode] -> special-node table or -1
    HkSpecialNode* spec;
    double* Q0;           // [nNode*3] internal force supplied through hk_upload_state (used once)
    // BC tables
    double* bc_value;     // per BC list entry
    int* bc_amp;          // per BC list entry -> amplitude table id or -1
    HkAmpTable* amp_tab;
    double* amp_time;
    double* amp_value;
    // elements
    int* conn;            // [8][nEp] 0-based node ids
    unsigned char* flag;  // 1 live, 3 deleted this step (state not yet zeroed), 0 deleted in the last step run (Qe still
                          // valid), 2 deleted and Qe cleared
    unsigned short* mat;  // 0-based material id
    HkMaterialDev* mats;
    double* ips;          // ip state, tile-blocked: [tile][gauss point 8][row 14][TL] with rows 0-5 stress, 6-11 strain,
                          // 12 eps (integ_eq_plastic_strain), 13 yield (integ_yield_stress): the 14 rows x TL
                          // elements of one Gauss point of one tile are ONE contiguous burst (14*TL*8 bytes)
    int element_mode;     // hk_params.element_mode (1: reference-order kernel)
    const long long* t_dev;   // step replay by CUDA graph: the step number t lives on the device (NULL: passed by value)
    int variant;          // element-kernel variant (hk_element.cu: kVariants; 1 = simple kernel)
    int n_sm;             // multiprocessors of the engine's device (grid of the persistent kernels)
    int TL;               // layout tile = elements per tile of the element kernel in use; nEp % TL == 0
    double* triax;        // [8][nEp]   integ_triax_stress (written on request)
    double* Qe;           // [24][nEp]
    // deletion log
    int* del_count;       // entries in del_list (all steps so far)
    int* del_block;       // deletion pass: marks per block of 1024 elements
    int* del_fresh;       // deletion pass: elements deleted in the step that just ran
    long long* del_list;  // (step << 32 | element) entries
    int del_cap;
    unsigned long long* counters;  // [0] negative jacobians [1] contact hits [2] contact tests [3] fixed-point overflow
    // contact force accumulators (128-bit fixed point): per slot 3 x {lo, hi}
    unsigned long long* cacc;
    double* halo_recv;    // [n_halo*3]
};

// index of (row, gauss point k, element e) in HkDev::ips
HK_HD long long hk_ip(const HkDev& d, int row, int k, long long e) {
    const long long t = e / d.TL;
    const long long j = e - t * d.TL;
    return ((t * 8 + k) * 14 + row) * d.TL + j;
}

struct HkPairDyn {               // current sizes of a pair's lists: device-resident, because exposed faces are
    int nn_i, nn_j, nTri;        // appended on the device (erode_element) without the host knowing
    int n_bucket;                // power of two >= 2*nn_i (<= cap_bucket)
    int n_cand;                  // per step: master triangles that survived the culls (hk_contact_cull_kernel)
};

struct HkPairDev {               // one ordered contact pair (ContactTriangle, J2:72-78)
    int i_instance, j_instance;  // 1-based instance ids (J2:272-354)
    int self;                    // i_instance == j_instance
    int cap_i, cap_j, cap_tri, cap_bucket;   // allocated lengths (== initial sizes when the surface cannot erode)
    int* nodes_i;                // 0-based node ids (c_nodes_i)
    int* nodes_j;
    int* t0; int* t1; int* t2;   // c_triangles columns, 0-based node ids
    int* tele;                   // c_triangles_eleid, 0-based
    double young;
    HkPairDyn* dyn;
    unsigned char* in_i;         // [nNode] membership of c_nodes_i / c_nodes_j (erodible pairs only, else NULL)
    unsigned char* in_j;
    // per-step work
    unsigned long long* bbox;    // 12 order-encoded doubles: min_i[3] max_i[3] min_j[3] max_j[3]
    int* cell_i;                 // [3][cap_i] cell coordinates of the i nodes
    int* head;                   // [cap_bucket] bucket heads (linked lists), -1 = empty
    int* next;                   // [cap_i]
    int* cand;                   // [cap_tri] ids of the step's surviving master triangles
};

// exposed-face update on the device (A10, J2:767-804 + add_surface_triangle J2:2167-2245): static per-instance tables
struct HkInstDev {
    long long element_offset, nElement;   // engine elements [element_offset, element_offset + nElement)
    int* surf;                   // [4][F] oriented face nodes (J2:1946-1992) as engine node ids (0-based), F = 6*nElement
    int* feleid;                 // [F] engine element (0-based) of every face
    int* twin;                   // [F] first face (in face-id order) of ANOTHER element with the same node set, or -1
};

struct HkErodeDev {              // everything erode_element needs
    int n_inst, n_pair;
    HkInstDev* inst;
    HkPairDev* pairs;            // device copy of the pair descriptors
    unsigned short* einst;       // [nElement] 1-based instance of every element (0: none)
    int* n_spec;                 // special-node table length / capacity, contact slots / capacity
    int* n_slots;
    int spec_cap, slot_cap;
    int* overflow;               // set when a capacity would be exceeded (checked at hk_sync)
    // partitioned meshes (hk_comm_erosion): the instance tables cover the GLOBAL mesh — `surf` holds engine-local node ids
    // (every node of an instance in contact is held or ghosted), `feleid` the engine-local element or -1 when another
    // rank owns it (that rank adds the triangle), `einst` is indexed by the global element; node_key[local node] = its
    // global id, the order in which `unique(sort(tri))` (J2:2242) lists an element's exposed nodes.  NULL: single domain
    const int* node_key;
};

struct HkContactParams {
    double d_lim, myu, kc_o, kc_s, cr_o, cr_s, ddiv_o, ddiv_s, d_time;
    int lsb_exp;                 // fixed-point LSB = 2^lsb_exp
    // v0.0.1 penetration-rate clamp (hk_params.contact_dmax_clamp; J1:2756, 2898, 618): per contact slot the largest
    // (clamped) penetration of this / the previous step, and d_max = max_n |d_disp_n| of the previous step as the bit
    // pattern of a non-negative double (so that an integer atomicMax orders it)
    int clamp;
    double* dnode;
    const double* dnode_pre;
    const unsigned long long* dmax;
};

// launchers (hk_exact.cu: built with -fmad=false; hk_element.cu: FMA allowed)
void hk_launch_nodal(const HkDev& d, double current_time, double d_time, double dt2, double dt2p,
                     int lsb_exp, int contact_on, int use_Q0, int mode, const int* list, long long n_list,
                     unsigned long long* dmax_out, cudaStream_t s);   // dmax_out: NULL or max |d_disp| accumulator
int hk_launch_element(const HkDev& d, long long step, int write_triax, cudaStream_t s);   // 0 or a CUDA error code
void hk_launch_contact(const HkDev& d, const HkPairDev& p, const HkContactParams& cp, cudaStream_t s);
// deletion pass of a step: elements the element kernel marked (flag 3) are appended to the deletion log in ascending
// id order (the reference's order, J2:701-735), their stress/strain zeroed (J2:742-756), and the faces they expose join
// the contact surfaces (er != NULL)
void hk_launch_step_set(long long* t_dev, long long t, cudaStream_t s);      // *t_dev = t
void hk_launch_step_advance(long long* t_dev, cudaStream_t s);               // ++*t_dev (last node of a captured step)
void hk_launch_deletion_pass(const HkDev& d, const HkErodeDev* er, long long step, cudaStream_t s, long long* n_launch,
                             long long* send = nullptr, const int* e_l2g = nullptr, int cap = 0);
// partitioned meshes: replays the deletions of ALL ranks of one step, `gathered` = [world][cap + 1] = {n, global 0-based
// element ids ascending} per rank in rank order (= ascending global id: ranks own contiguous element blocks)
void hk_launch_erode_replay(const HkDev& d, const HkErodeDev& E, const long long* gathered, int world, int cap, cudaStream_t s);
void hk_launch_cacc_zero(const HkDev& d, const int* n_slots, int slot_cap, cudaStream_t s);
void hk_launch_velo_from_rec(const HkDev& d, double d_time, cudaStream_t s);
void hk_launch_gather_Q(const HkDev& d, double* Q_out, cudaStream_t s);
void hk_launch_nodes_export(const HkDev& d, const int* nodes, long long n, double* out, cudaStream_t s);
void hk_launch_nodes_import(const HkDev& d, const int* nodes, const long long* src, long long n, const double* in, cudaStream_t s);
void hk_launch_cacc_export(const HkDev& d, const int* nodes, long long n, unsigned long long* out, cudaStream_t s);
void hk_launch_cacc_import(const HkDev& d, const int* nodes, long long n, const unsigned long long* in, long long n_ranks, cudaStream_t s);
void hk_launch_state_export(const HkDev& d, const int* nodes, long long n, double* out, cudaStream_t s);
void hk_launch_state_import(const HkDev& d, const int* nodes, long long n, const double* in, double d_time, cudaStream_t s);
void hk_launch_cacc_export_limbs(const HkDev& d, const int* nodes, long long n, long long* out, cudaStream_t s);
void hk_launch_cacc_import_limbs(const HkDev& d, const int* nodes, long long n, const long long* in, cudaStream_t s);
// halo (multi-GPU interface nodes): own partial forces, per-neighbour send blocks, rank-ordered total in d.halo_recv
void hk_launch_halo_pack(const HkDev& d, const int* nodes, long long n, double* out, const double* Q0, cudaStream_t s);
void hk_launch_halo_gather(const double* own, const int* slots, long long n, double* send, cudaStream_t s);
void hk_launch_halo_accumulate(const HkDev& d, const int* slots, long long n, const double* recv, cudaStream_t s);
void hk_launch_triax(const HkDev& d, cudaStream_t s);
void hk_launch_element_volume(const HkDev& d, double* V_out, cudaStream_t s);
// layout transposes between the reference's AoS (6,nip)/(nip) arrays and the SoA rows
// rows [row0, row0+ncomp) of the ip state <-> the reference's (ncomp, nip) array, element range [e0, e0+ne)
void hk_launch_ip_to_dev(const double* aos, const HkDev& d, int row0, int ncomp, long long e0, long long ne, cudaStream_t s);
void hk_launch_ip_to_aos(const HkDev& d, double* aos, int row0, int ncomp, long long e0, long long ne, cudaStream_t s);
// triax [8][nEp] -> (nip) of the reference
void hk_launch_triax_to_aos(const HkDev& d, double* aos, long long e0, long long ne, cudaStream_t s);
void hk_launch_element_means(const HkDev& d, double* emean, cudaStream_t s);                 // [14][nEp]
void hk_launch_node_means(const HkDev& d, const double* emean, double* out, int raw, cudaStream_t s);   // [16][nNode]
void hk_launch_state_summary(const HkDev& d, unsigned long long* out4, cudaStream_t s);   // out4 preset {0,~0,0,0}
double hk_decode_double(unsigned long long order_encoded);
int hk_upload_pusai(const double* P);
void hk_launch_element_exact(const HkDev& d, long long step, int write_triax, cudaStream_t s);
int hk_element_variant_from_env();        // HK_ELEMENT_VARIANT / HK_ELEMENT_KERNEL (A/B switches), else the default
long long hk_element_tile(int variant);   // nEp must be a multiple of this
void hk_launch_external_force(const HkDev& d, double* F_out, int lsb_exp, int contact_on, cudaStream_t s);
