/* libwgs_b200 — C ABI of the B200-native WarpedGANSpace hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name starts with `h_`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); no call synchronises the
 *     host unless documented;
 *   - return value 0 = success, non-zero = failure, message via wgs_last_error();
 *   - tensors are dense row-major; activations are NHWC ("channels last") fp32 unless stated.
 *
 * Each entry point cites the reference interface it replaces (paths under chi0tzp/WarpedGANSpace).
 */
#ifndef WGS_B200_H
#define WGS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------------------------------- */
const char* wgs_last_error(void);
int  wgs_version(void);
unsigned long long wgs_launch_count(void);       /* kernels launched through this library so far */
void wgs_reset_launch_count(void);
int  wgs_device_info(int* sms, int* cc);         /* fails unless the current device is sm_100 */

/* ---- SupportSets RBF warp --------------------------------------------------------------------- *
 * Replaces SupportSets.forward, lib/support_sets.py:81-101 (and its autograd backward), with the
 * one-hot mask given as row indices:  out[b] = mag[b] * grad_f(z[b]) / ||grad_f(z[b])||,
 *   grad_f = -2 * sum_j alpha[k,j] * gamma_k * exp(-gamma_k * |z - s_kj|^2) * (z - s_kj),  k = idx[b].
 * support_sets [K, n_vec*d] (n_vec = 2 * num_support_dipoles), alphas [K, n_vec], loggamma [K] or NULL
 * (then gamma = fixed_gamma, the learn_gammas=False branch :93), idx [B] int64, z [B, d],
 * mag [B] or NULL (= 1, the bare module output), out [B, d].  d % 4 == 0, d <= 1024.             */
int wgs_rbf_warp_forward(const float* support_sets, const float* alphas, const float* loggamma,
                         float fixed_gamma, const long long* idx, const float* z, const float* mag,
                         float* out, int B, int K, int n_vec, int d, void* stream);

/* Backward of the above for upstream gradient dout [B, d].  d_support_sets [K, n_vec*d] and
 * d_loggamma [K] / d_alphas [K, n_vec] are ACCUMULATED into (atomics; zero them first; only the
 * rows named by idx are touched, as in the reference's one-hot matmul backward); dz [B, d] is
 * overwritten.  Any of the four outputs may be NULL.                                             */
int wgs_rbf_warp_backward(const float* support_sets, const float* alphas, const float* loggamma,
                          float fixed_gamma, const long long* idx, const float* z, const float* mag,
                          const float* dout, float* d_support_sets, float* d_loggamma, float* d_alphas,
                          float* dz, int B, int K, int n_vec, int d, void* stream);

/* Traversal chains, traverse_latent_space.py:369-438: for each chain c, `steps` sequential steps
 * shift = +-eps * S(path[c], code); code += shift, in both directions from start[c].
 * codes, shifts: [chains, 2*steps+1, d], most negative step first, centre frame = (start, 0).     */
int wgs_rbf_traverse(const float* support_sets, const float* alphas, const float* loggamma,
                     float fixed_gamma, const long long* path, const float* start, float eps, int steps,
                     float* codes, float* shifts, int chains, int K, int n_vec, int d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WGS_B200_H */
