// kernels.h -- host-callable launchers of the sm_100a kernels (definitions in kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tile_core.cuh"

namespace dvd {

constexpr int BLK_BITS = 10;  // sampler / reduction leaf block: 2^10 amplitudes

cudaError_t kernels_init();   // one-time function attributes (dynamic shared memory opt-in)
long ring_min_tiles(int sm_count);   // fewest tiles of a dense pass for the two-group persistent form (DVD_RING_MIN_TILES)

// Tiled multi-gate pass (n_local >= TILE_BITS).  pp (description + op list) travels as the kernel parameter.
cudaError_t launch_tile_pass(cplx* amp, PassParams& pp, cudaStream_t s);   // fills pp.pd.last_switch

// One gate, one pass, any size (used when n_local < TILE_BITS, and as the un-fused debug path).
struct SimpleOp { double m[8]; int32_t tbit; int32_t cbit; };   // cbit < 0: no control
cudaError_t launch_simple_gate(cplx* amp, int n_local, uint64_t rank_bits, const SimpleOp& op, cudaStream_t s);

cudaError_t launch_set_basis_state(cplx* amp, uint64_t index, double value, cudaStream_t s);
// amp[i] = 0 for every i with (i & zmask) != 0: turns the implied zeros of support tracking into stored zeros
cudaError_t launch_zero_outside_support(cplx* amp, int n_local, uint64_t zmask, cudaStream_t s);

// probs[i] = re^2 + im^2 for i in [first, first+count)
cudaError_t launch_probabilities(const cplx* amp, uint64_t first, uint64_t count, double* probs, cudaStream_t s);

// Pairwise summation tree over |amp|^2.  `tree` receives the levels >= nb (nb = min(BLK_BITS,n_local)):
// level l (l = nb..n_local) starts at tree_level_offset(n_local, l) and has 2^(n_local-l) entries.
uint64_t tree_level_offset(int n_local, int level);
uint64_t tree_size(int n_local);
cudaError_t launch_build_tree(const cplx* amp, int n_local, double* tree, cudaStream_t s);

// Tree-descent sampler: out[s] = local index for xsi = u[s] * total (total = tree root).
// sel == nullptr: all shots; else shot s is processed only if sel[s] == sel_value.
cudaError_t launch_sample(const cplx* amp, int n_local, const double* tree, const double* u,
                          const int32_t* sel, int32_t sel_value, uint64_t index_offset,
                          uint64_t shots, unsigned long long* out, cudaStream_t s);

// Reference-order sampler (strict left-to-right cumulative sums, utils.rs:258-277): block_cum[b] = cum[(b + 1) * 2^nb]
// (nb = min(BLK_BITS, n_local); 2^(n_local - nb) entries), built by one warp in index order; then one thread per shot.
cudaError_t launch_seq_block_cum(const cplx* amp, int n_local, double* block_cum, cudaStream_t s);
cudaError_t launch_sample_seq(const cplx* amp, int n_local, const double* block_cum, const double* u, const int32_t* sel,
                              int32_t sel_value, uint64_t index_offset, uint64_t shots, unsigned long long* out, cudaStream_t s);

// out[s*n_obs + o] = ((samples[s] >> qubits[o]) & 1) ? -1 : +1
cudaError_t launch_extract_expectation(const unsigned long long* samples, uint64_t shots, const int* qubits,
                                       int n_obs, double* out, cudaStream_t s);

// Exact <Z_q> numerators: out[q] = sum_i p_i * (1 - 2*bit_q(i)) over the local chunk, q < n_total
// (bits >= n_local are taken from rank_bits).  partial is scratch of ez_partial_size() doubles.
uint64_t ez_partial_size();
cudaError_t launch_expectation_z(const cplx* amp, int n_local, int n_total, uint64_t rank_bits,
                                 double* partial, double* out, cudaStream_t s);

// Global<->local qubit swap staging: gather / scatter the half of the chunk whose bit `lq` equals
// `bitval`, elements [first, first+count) of that half, to / from a contiguous buffer.
cudaError_t launch_pack_half(const cplx* amp, int lq, int bitval, uint64_t first, uint64_t count, cplx* buf, cudaStream_t s);
cudaError_t launch_unpack_half(cplx* amp, int lq, int bitval, uint64_t first, uint64_t count, const cplx* buf, cudaStream_t s);

// Direct NVLink swap: exchange elements [e_begin, e_end) of this rank's leaving half (bit lq == 1 - my_bit)
// with the partner's leaving half (its bit lq == my_bit); `peer` is the partner's state mapped with CUDA IPC.
cudaError_t launch_swap_peer(cplx* mine, cplx* peer, int lq, int my_bit, uint64_t e_begin, uint64_t e_end, cudaStream_t s);

// interleaved complex128 <-> separate real / imaginary arrays (device side of dvd_read_state / dvd_load_state)
cudaError_t launch_split_re_im(const cplx* amp, uint64_t count, double* re, double* im, cudaStream_t s);
cudaError_t launch_join_re_im(cplx* amp, uint64_t count, const double* re, const double* im, cudaStream_t s);

// <a|b> partial dot (conj(a).b), deterministic two-stage reduction; out[0]=re, out[1]=im
cudaError_t launch_dot(const cplx* a, const cplx* b, uint64_t count, double* partial, double* out, cudaStream_t s);

}  // namespace dvd
