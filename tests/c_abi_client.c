/* A plain C client of include/fcfv_b200.h (what a non-Python host links against): takes the address of every entry
 * point and, when a GPU is present, runs the smallest meaningful sequence (one triangle, all faces Dirichlet: Kuu is
 * the 6x6 identity).  Built and linked by tests/test_cabi.py with gcc; exit code 0 = OK, 77 = no CUDA device. */
#include <stdio.h>
#include <string.h>
#include "../include/fcfv_b200.h"

int main(void) {
    void* syms[] = {(void*)fcfv_create, (void*)fcfv_destroy, (void*)fcfv_last_error, (void*)fcfv_version, (void*)fcfv_set_mesh,
                    (void*)fcfv_update_fields, (void*)fcfv_set_ownership, (void*)fcfv_geometry, (void*)fcfv_compute_local,
                    (void*)fcfv_assemble, (void*)fcfv_schur, (void*)fcfv_run_assemble_schur, (void*)fcfv_index_width,
                    (void*)fcfv_block_size, (void*)fcfv_export, (void*)fcfv_export_rhs, (void*)fcfv_device_csr, (void*)fcfv_device_rhs,
                    (void*)fcfv_residuals, (void*)fcfv_element_values, (void*)fcfv_ph_residual, (void*)fcfv_ph_update_p,
                    (void*)fcfv_ph_fusc, (void*)fcfv_center_pressure, (void*)fcfv_run_local, (void*)fcfv_run_assemble,
                    (void*)fcfv_run_residuals, (void*)fcfv_run_pipeline, (void*)fcfv_generate_structured, (void*)fcfv_comm_init,
                    (void*)fcfv_compute_fcfv, (void*)fcfv_sparse_face_data, (void*)fcfv_sparse_face_info, (void*)fcfv_debug_bounds_violations};
    for (unsigned k = 0; k < sizeof syms / sizeof syms[0]; k++)
        if (!syms[k]) return 2;
    fcfv_t* h = NULL;
    if (fcfv_create(&h, 0) != FCFV_OK) { printf("no CUDA device: link check only\n"); return 77; }
    /* one triangle (0,0) (1,0) (0,1); faces 1,2,3 = edges (v1,v2), (v2,v3), (v3,v1); everything Dirichlet */
    const int64_t e2f[3] = {1, 2, 3}, bc[3] = {1, 1, 1};
    const double Om[1] = {0.5}, G[3] = {1.0, 1.4142135623730951, 1.0};
    const double nx[3] = {0.0, 0.7071067811865476, -1.0}, ny[3] = {-1.0, 0.7071067811865476, 0.0};
    const double ke[1] = {1.0}, dl[1] = {1.0}, taue[1] = {1.0}, zero1[1] = {0.0}, zero3[3] = {0.0, 0.0, 0.0};
    const double Vx[3] = {1.0, 2.0, 3.0}, Vy[3] = {-1.0, -2.0, -3.0};
    double alpha[1], beta[2], Z[4], tau[3], fu[6], fp[1], val[6];
    int64_t nuu = 0, nup = 0, nmu = 0, colptr[7], rowval[6];
    double sec = 0.0;
    /* ComputeFCFV(mesh, ...) as the shims issue it: one call that takes the mesh with it ... */
    double alpha1[1], beta1[2], Z1[4], tau1[3];
    int rc = fcfv_compute_fcfv(h, 1, 3, e2f, bc, Om, nx, ny, G, ke, dl, taue, 1, NULL, FCFV_GRADIENT, 1.0, zero1, zero1, Vx, Vy, zero3, zero3,
                               zero3, zero3, alpha1, beta1, Z1, tau1);
    /* ... and the separate calls it stands for */
    if (!rc) rc = fcfv_set_mesh(h, 1, 3, e2f, bc, Om, nx, ny, G, ke, dl, taue, 1);
    if (!rc) rc = fcfv_compute_local(h, FCFV_GRADIENT, 1.0, zero1, zero1, Vx, Vy, alpha, beta, Z, tau);
    if (!rc && !(alpha1[0] == alpha[0] && beta1[0] == beta[0] && beta1[1] == beta[1] && Z1[0] == Z[0] && Z1[3] == Z[3] && tau1[2] == tau[2])) {
        printf("fcfv_compute_fcfv differs from fcfv_set_mesh + fcfv_compute_local\n"); fcfv_destroy(h); return 1;
    }
    if (!rc) rc = fcfv_assemble(h, FCFV_LOOP_OLD, FCFV_GRADIENT, 1.0, NULL, NULL, NULL, Vx, Vy, zero3, zero3, zero3, zero3, 0, &nuu, &nup, &nmu, &sec);
    if (!rc) rc = fcfv_export(h, FCFV_KUU, FCFV_CSC1_INT64, colptr, rowval, val);
    if (!rc) rc = fcfv_export_rhs(h, fu, fp);
    if (rc) { printf("error %d: %s\n", rc, fcfv_last_error(h)); fcfv_destroy(h); return 1; }
    int ok = nuu == 6 && nup == 0;
    for (int k = 0; k < 6 && ok; k++) ok = colptr[k] == k + 1 && rowval[k] == k + 1 && val[k] == 1.0;
    for (int k = 0; k < 3 && ok; k++) ok = fu[k] == Vx[k] && fu[3 + k] == Vy[k];
    printf("%s: nnz(Kuu)=%lld alpha=%g tsparse=%.2e s\n", ok ? "C client OK" : "C client MISMATCH", (long long)nuu, alpha[0], sec);
    fcfv_destroy(h);
    return ok ? 0 : 1;
}
