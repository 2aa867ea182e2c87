/* tests only: route the driver's calls to the host-compiled kernel build (tests/emu/libhakai_emu.so, prefix hke_) so
 * that the C program's own logic can be exercised in the GPU-less container. */
#define hk_default_params hke_default_params
#define hk_create hke_create
#define hk_destroy hke_destroy
#define hk_last_error hke_last_error
#define hk_set_mesh hke_set_mesh
#define hk_add_material hke_add_material
#define hk_add_bc hke_add_bc
#define hk_add_ic hke_add_ic
#define hk_finalize hke_finalize
#define hk_step hke_step
#define hk_download hke_download
#define hk_node_output hke_node_output
