// hk_setup.cu — contact set-up on the device (A12): the faces of an instance, their outward orientation, which of them are
// exterior and, for every face, the face a deletion of its element would expose.
//
// Replaces get_element_face (J2:1946-1992) and the face matching of get_surface_triangle (J2:1996-2084) /
// add_surface_triangle (J2:2167-2245).  The reference matches every face against every other face of the instance,
// O(F^2) — what stops it from running an 8 M-element deck at all; here faces are radix-sorted by their sorted node
// 4-tuple (cub::DeviceRadixSort, four stable 32-bit passes, face id as payload), so faces with the same node set are
// adjacent, in ascending face id.  From the sorted order one kernel derives, per group of equal tuples,
//   * the faces the reference's pairing loop would emit as exterior (J2:2040-2084): it pairs the first unmatched face
//     with the next equal one, so a group of odd size emits its last member — and never the very last face of the
//     instance (loop bound 1 : 6nE-1, J2:2040);
//   * twin[f]: the first face, in face-id order, of ANOTHER element with the same node set (what add_surface_triangle
//     finds for face f of a deleted element), or -1.
// The sort is library code (CUB); it runs once per instance at set-up, not on the per-step path.
// The host-compiled debugging build (-DHK_EMU) does the same with std::sort.
#include <algorithm>
#include <array>
#include <vector>

#include "hk_common.h"

#ifndef HK_EMU
#include <cub/cub.cuh>
#endif

namespace {

// local node ids of the 6 faces (J2:1951-1958) and the flipped order [0,3,2,1] (J2:1985-1989)
HK_HD void face_nodes(const int n8[8], int j, int out[4]) {
    const int loc[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}};
    for (int a = 0; a < 4; ++a) out[a] = n8[loc[j][a]];
}

// one element: its 6 oriented faces (part-local 1-based node ids) and their sorted tuples
HK_HD void element_faces(const int* conn, long long nEp, const double* X, long long node_offset, long long e_global,
                         long long e_local, long long F, int* surf /*[4][F]*/, unsigned* key /*[4][F]*/) {
    int n8[8];
    double ctr[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < 8; ++a) {
        n8[a] = conn[(long long)a * nEp + e_global];                 // engine node id, 0-based
        for (int c = 0; c < 3; ++c) ctr[c] += X[3ll * n8[a] + c];
    }
    for (int c = 0; c < 3; ++c) ctr[c] = ctr[c] / 8;
    for (int j = 0; j < 6; ++j) {
        int f[4];
        face_nodes(n8, j, f);
        const double* p1 = X + 3ll * f[0];
        const double* p2 = X + 3ll * f[1];
        const double* p4 = X + 3ll * f[3];
        const double v1[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        const double v2[3] = {p4[0] - p1[0], p4[1] - p1[1], p4[2] - p1[2]};
        const double nv[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
        const double vc[3] = {ctr[0] - p1[0], ctr[1] - p1[1], ctr[2] - p1[2]};
        if (nv[0] * vc[0] + nv[1] * vc[1] + nv[2] * vc[2] > 0.0) { const int t = f[1]; f[1] = f[3]; f[3] = t; }   // J2:1985-1989
        const long long fid = 6 * e_local + j;
        unsigned k[4];
        for (int a = 0; a < 4; ++a) {
            const int pl = (int)(f[a] - node_offset) + 1;               // part-local 1-based
            surf[(long long)a * F + fid] = pl;
            k[a] = (unsigned)pl;
        }
        // sort the 4 ids (5-comparator network)
#define HK_CSWAP(i, j_) { if (k[i] > k[j_]) { const unsigned t = k[i]; k[i] = k[j_]; k[j_] = t; } }
        HK_CSWAP(0, 1) HK_CSWAP(2, 3) HK_CSWAP(0, 2) HK_CSWAP(1, 3) HK_CSWAP(1, 2)
#undef HK_CSWAP
        for (int a = 0; a < 4; ++a) key[(long long)a * F + fid] = k[a];
    }
}

// position p of the sorted order: exterior flag and twin of face perm[p]
HK_HD void group_resolve(const unsigned* key /*[4][F]*/, const int* perm, long long F, long long p, unsigned char* exterior,
                         int* twin) {
    const int f = perm[p];
    auto same = [&](int a, int b) {
        return key[a] == key[b] && key[F + a] == key[F + b] && key[2 * F + a] == key[2 * F + b] && key[3 * F + a] == key[3 * F + b];
    };
    long long lo = p, hi = p;
    while (lo > 0 && same(perm[lo - 1], f)) --lo;
    while (hi + 1 < F && same(perm[hi + 1], f)) ++hi;
    const long long size = hi - lo + 1, rank = p - lo;
    exterior[f] = ((size & 1) && rank == size - 1 && f != F - 1) ? 1 : 0;
    int t = -1;
    for (long long q = lo; q <= hi; ++q)                             // ascending face id inside the group
        if (perm[q] / 6 != f / 6) { t = perm[q]; break; }
    twin[f] = t;
}

#ifndef HK_EMU
__global__ void hk_faces_kernel(const int* conn, long long nEp, const double* X, long long node_offset, long long element_offset,
                                long long nElement, int* surf, unsigned* key) {
    const long long el = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (el < nElement) element_faces(conn, nEp, X, node_offset, element_offset + el, el, 6 * nElement, surf, key);
}
__global__ void hk_gather_key_kernel(const unsigned* key_col, const int* perm, unsigned* out, long long F) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < F) out[i] = key_col[perm[i]];
}
__global__ void hk_iota_kernel(int* p, long long F) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < F) p[i] = (int)i;
}
__global__ void hk_group_kernel(const unsigned* key, const int* perm, long long F, unsigned char* exterior, int* twin) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p < F) group_resolve(key, perm, F, p, exterior, twin);
}
#endif

}  // namespace

// Faces of one instance.  d_conn [8][nEp] / d_X [3*nNode]: device copies of the mesh (engine node ids, 0-based).
// Outputs (host): surf (F,4) column-major part-local 1-based, twin (F), exterior face ids ascending.
// Returns 0 or a CUDA error code.
int hk_setup_instance_faces(const int* d_conn, long long nEp, const double* d_X, long long node_offset, long long element_offset,
                            long long nElement, cudaStream_t s, std::vector<int>& surf, std::vector<int>& twin,
                            std::vector<int>& exterior_ids) {
    const long long F = 6 * nElement;
    surf.assign((size_t)4 * F, 0);
    twin.assign((size_t)F, -1);
    exterior_ids.clear();
    if (F == 0) return 0;
    std::vector<unsigned char> ext((size_t)F, 0);
#ifndef HK_EMU
    int* p_surf = nullptr; unsigned* p_key = nullptr; int *p_perm = nullptr, *p_perm2 = nullptr, *p_twin = nullptr;
    unsigned *p_k = nullptr, *p_k2 = nullptr; unsigned char* p_ext = nullptr; void* p_tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t rc = cudaSuccess;
    auto done = [&](cudaError_t r) {
        cudaFree(p_surf); cudaFree(p_key); cudaFree(p_perm); cudaFree(p_perm2); cudaFree(p_twin); cudaFree(p_k); cudaFree(p_k2);
        cudaFree(p_ext); cudaFree(p_tmp);
        return (int)r;
    };
#define HK_TRY(x) do { rc = (x); if (rc != cudaSuccess) return done(rc); } while (0)
    HK_TRY(cudaMalloc(&p_surf, (size_t)4 * F * sizeof(int)));
    HK_TRY(cudaMalloc(&p_key, (size_t)4 * F * sizeof(unsigned)));
    HK_TRY(cudaMalloc(&p_perm, (size_t)F * sizeof(int)));
    HK_TRY(cudaMalloc(&p_perm2, (size_t)F * sizeof(int)));
    HK_TRY(cudaMalloc(&p_twin, (size_t)F * sizeof(int)));
    HK_TRY(cudaMalloc(&p_k, (size_t)F * sizeof(unsigned)));
    HK_TRY(cudaMalloc(&p_k2, (size_t)F * sizeof(unsigned)));
    HK_TRY(cudaMalloc(&p_ext, (size_t)F));
    const int B = 256;
    const unsigned gE = (unsigned)((nElement + B - 1) / B), gF = (unsigned)((F + B - 1) / B);
    hk_faces_kernel<<<gE, B, 0, s>>>(d_conn, nEp, d_X, node_offset, element_offset, nElement, p_surf, p_key);
    hk_iota_kernel<<<gF, B, 0, s>>>(p_perm, F);
    HK_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, p_k, p_k2, p_perm, p_perm2, (int)F, 0, 32, s));
    HK_TRY(cudaMalloc(&p_tmp, tmp_bytes));
    for (int comp = 3; comp >= 0; --comp) {          // LSD: least significant tuple component first, every pass stable
        hk_gather_key_kernel<<<gF, B, 0, s>>>(p_key + (size_t)comp * F, p_perm, p_k, F);
        HK_TRY(cub::DeviceRadixSort::SortPairs(p_tmp, tmp_bytes, p_k, p_k2, p_perm, p_perm2, (int)F, 0, 32, s));
        std::swap(p_perm, p_perm2);
    }
    hk_group_kernel<<<gF, B, 0, s>>>(p_key, p_perm, F, p_ext, p_twin);
    HK_TRY(cudaMemcpyAsync(surf.data(), p_surf, (size_t)4 * F * sizeof(int), cudaMemcpyDeviceToHost, s));
    HK_TRY(cudaMemcpyAsync(twin.data(), p_twin, (size_t)F * sizeof(int), cudaMemcpyDeviceToHost, s));
    HK_TRY(cudaMemcpyAsync(ext.data(), p_ext, (size_t)F, cudaMemcpyDeviceToHost, s));
    HK_TRY(cudaStreamSynchronize(s));
    HK_TRY(cudaGetLastError());
#undef HK_TRY
    done(cudaSuccess);
#else
    (void)s;
    std::vector<unsigned> key((size_t)4 * F);
    for (long long el = 0; el < nElement; ++el)
        element_faces(d_conn, nEp, d_X, node_offset, element_offset + el, el, F, surf.data(), key.data());
    std::vector<int> perm((size_t)F);
    for (long long i = 0; i < F; ++i) perm[i] = (int)i;
    std::sort(perm.begin(), perm.end(), [&](int a, int b) {
        for (int c = 0; c < 4; ++c)
            if (key[(size_t)c * F + a] != key[(size_t)c * F + b]) return key[(size_t)c * F + a] < key[(size_t)c * F + b];
        return a < b;
    });
    for (long long p = 0; p < F; ++p) group_resolve(key.data(), perm.data(), F, p, ext.data(), twin.data());
#endif
    for (long long f = 0; f < F; ++f)
        if (ext[f]) exterior_ids.push_back((int)f);
    return 0;
}
