// nagp_kernels.cuh — launch interfaces shared by the C-ABI layer and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nagp_tree.cuh"

namespace nagp {

// One fused Gram -> Cholesky -> solve problem per instance b in [0, B). Instance b belongs to
// scenario s = b / P and particle p = b % P (P = B for plain logML batches).
struct FusedArgs {
    int64_t B;
    int64_t P;
    const uint8_t *prog;
    const int64_t *prog_off;     // [P+1]
    const double *theta;
    const int64_t *theta_off;    // [P+1]
    int64_t theta_stride_k;      // doubles between scenarios (0 = shared)
    const double *noise;         // [P] or [K,P]
    int64_t noise_stride_k;
    double jitter;
    double noise_pred;           // < 0: instance noise on forecast block
    int n, k, h;                 // q = n + k + h
    const double *t;             // [q]
    const int32_t *g;            // [q] or null
    double step;
    int G;                       // max lag + 1 (lag-grid mode), else 0
    const double *y1;            // [n] (+ b * y1_stride)
    int64_t y1_stride;           // per-instance stride of y1 (0 = shared)
    const double *y2;            // [K,k] or null: scenario values, indexed by s
    double ya, yb;
    // outputs (all nullable except info)
    double *logml_n;             // [B]
    double *logml_m;             // [B] (needs y2 or k == 0)
    const double *logw0;         // [P] nullable
    double *logw;                // [B] = logw0[p] + logml_m - logml_n
    double *mu;                  // [B,h] original units (needs y for all m rows)
    double *L33;                 // [B,h,h] row-major lower / ya
    double *proj;                // [B,k+h]: sum_{j<n} L[n+r][j] z1[j]   (scenario-shared fast path)
    double *Ltail;               // [B,k+h,k+h]: L[n+r][n+c] (scaled space, zeros above diagonal)
    int32_t *info;               // [B]
    int ntab_cap, ncp_cap;       // shared-memory table slots per CTA
    const TreeProgram *compiled; // [P] programs compiled on the host with these capacities (nullable: compile on device)
    double *Lkeep;               // nullable: the factor of every instance, tile-packed operand layout [B, tri(nt)*64]
    double *zkeep;               // nullable (with Lkeep): z = L^-1 y per instance [B, 8*nt]
    double *Wkeep;               // nullable (with Lkeep): inverses of the diagonal tiles, operand layout [B, nt*64]
};

size_t fused_smem_bytes_v1(int q, int G, int ntab_cap, int ncp_cap);
cudaError_t launch_fused_v1(const FusedArgs &a, cudaStream_t stream);

// Tile (DMMA) kernel: shared-memory plan, persistent grid size and launch.
struct V2Plan {
    int ok;
    int nt;
    int aux_off[5];
    int aux_smem[5];
    int scratch_stride;   // bytes of global scratch per CTA (0 when every table fits in shared memory)
    size_t smem_bytes;
};
int fused_v2_max_q();
V2Plan plan_fused_v2(int q, int G, int ntheta_cap, int ntab_cap, int ncp_cap, int smem_optin, int smem_per_sm);
int fused_v2_grid(const V2Plan &pl, int64_t B, int num_sms);
cudaError_t launch_fused_v2(const FusedArgs &a, const V2Plan &pl, char *scratch, unsigned long long *work_counter,
                            int grid, cudaStream_t stream);

// Slot kernel (nagp_fused_v3.cu): three matrices in flight per SM in one role-specialised CTA, factor in a pool of
// recycled tile slots, Gram on demand. Needs host-compiled programs whose stationary leaves are all tabulated.
struct V3Plan {
    int ok;
    int nt, ns, lead, iep;       // tile rows, pool slots, Gram lead (columns), first tile row kept for the epilogue
    int off_tt, off_gg, off_map, off_need, off_slots, slot_stride;
    int o_ctrl, o_tp, o_yv, o_diag, o_inv, o_th, o_tab, o_sig, o_pool;   // inside a slot's region
    size_t smem_bytes;
    unsigned char smap[232];     // tile (I, P) -> pool slot, at tri(I) + P (nt <= 21)
    unsigned char need[24];      // row-owner steps that must be complete before Gram column C may be written
};
int fused_v3_max_q();
V3Plan plan_fused_v3(int n, int k, int h, bool tail_rows_from_n, int G, int ntheta_cap, int ntab_cap, int ncp_cap,
                     int smem_optin);
int fused_v3_grid(int64_t B, int num_sms);
cudaError_t launch_fused_v3(const FusedArgs &a, const V3Plan &pl, unsigned long long *work_counter, int grid,
                            cudaStream_t stream);

// Large path (q beyond shared memory): factor in HBM as tile-packed operand-layout tiles.
struct LargePlan {
    int ok;
    int nt, ntp;          // tile rows of the problem / rounded up to a block of 8
    int ntp_cap;          // tile rows of the storage capacity (>= ntp)
    int yrow;             // tile row index of z = L^-1 y inside the storage (= ntp_cap)
    size_t L_stride;      // doubles per factor slot
    int aux_off[5];
    int aux_smem[5];
    int scratch_stride;
    int ring;             // rank-append only: shared memory holds the streaming ring
    size_t smem_bytes;
};
int large_max_q();
// ring: rank-append launches that touch a single tile row (the HBM-bound case) stream the factor through a shared-memory
// ring; block appends keep the space for the lag / sigma tables instead.
LargePlan plan_large(int q, int q_cap, int G, int ntheta_cap, int ntab_cap, int ncp_cap, int smem_per_sm, bool append,
                     bool ring = false);
int large_grid(const LargePlan &pl, int64_t B, int num_sms, bool append);
// keep = 1: instance b's factor goes to slot b of L (and its inverse diagonal tiles to W); 0: L is a
// per-CTA workspace of `grid` slots.
// dlogml != nullptr: block-append mode on kept factors (keep = 1): the tile rows from n_old / 8 on are (re)computed against
// the stored rows above them, the increment of logML over the points n_old .. a.n-1 goes to dlogml and is added to logml_acc.
cudaError_t launch_chol_large(const FusedArgs &a, const LargePlan &pl, char *scratch, double *L, int keep, double *W,
                              unsigned long long *work_counter, int grid, cudaStream_t stream, int n_old = 0,
                              double *dlogml = nullptr, double *logml_acc = nullptr);
// Extend the stored factors (slots 0..B-1) from n_old to a.n points in place; a.t/a.g/a.y1 cover all a.n points.
cudaError_t launch_rank_append(const FusedArgs &a, const LargePlan &pl, char *scratch, double *L, double *W,
                               int n_old, double *logml, double *dlogml, int grid, cudaStream_t stream);

// Gradient of the log marginal likelihood: consumes the factors kept by the tile kernel.
struct GradArgs {
    int64_t B, P;
    const uint8_t *prog; const int64_t *prog_off;
    const double *theta; const int64_t *theta_off; int64_t theta_stride_k;
    int n;                       // observed points (k = h = 0)
    const double *t; const int32_t *g; double step;
    const double *L;             // [B, tri(nt)*64] kept factors
    const double *z;             // [B, 8*nt]
    double *grad_theta;          // [B/P scenarios][theta total]: d logML / d theta slot
    double *grad_noise;          // [B]
    double *S;                   // nullable global scratch [grid][n*n] when n*n doubles do not fit shared memory
    const int32_t *info;         // [B] from the factorisation: instances with info != 0 get NaN gradients
    // tile version only
    const double *Winv;          // [B, nt*64] inverses of the diagonal tiles of L (FusedArgs::Wkeep)
    int G;                       // lag-grid extent (0: pairwise times)
    int ntab_cap, ncp_cap;       // table capacities the programs were compiled with
    const TreeProgram *compiled; // [P] nullable
};
size_t grad_smem_bytes(int n, int smem_optin, bool *s_in_smem);
cudaError_t launch_grad(const GradArgs &a, int grid, size_t smem_bytes, cudaStream_t stream);

// Tile (DMMA) version: K^-1 in place over the factor in shared memory, lag-binned reverse mode.
struct GradTilePlan {
    int ok;
    int nt, Gd, nsec;
    int region_bytes, scratch_stride;
    size_t smem_bytes;
};
GradTilePlan plan_grad_tile(int n, int G, int ntab_cap, int ncp_cap, int smem_optin, int smem_per_sm);
int grad_tile_grid(const GradTilePlan &pl, int64_t B, int num_sms);
cudaError_t launch_grad_tile(const GradArgs &a, const GradTilePlan &pl, char *scratch, unsigned long long *work_counter,
                             int grid, cudaStream_t stream);

// Scenario-shared fast path: per (scenario, particle) O(k^2 + hk) tail of the forward solve.
struct AppendArgs {
    int64_t K, P;
    int k, h;
    const double *y2;      // [K,k] scaled
    const double *proj;    // [P,k+h]
    const double *Ltail;   // [P,k+h,k+h]
    const double *logw0;   // [P] nullable
    const double *logml_n; // unused (kept for symmetry)
    double ya, yb;
    double *logw;          // [K,P]
    double *mu;            // [K,P,h] nullable
};
cudaError_t launch_append(const AppendArgs &a, cudaStream_t stream);

struct DrawArgs {
    int64_t K, P;
    int h;
    int64_t D;
    const double *logw;    // [K,P]
    const double *mu; int64_t mu_stride_k;
    const double *L; int64_t l_stride_k;
    const int32_t *comp;   // [K,D] nullable
    const double *u;       // [K,D] nullable
    const double *u_res;   // [K,P] nullable
    double ess_thr;
    const double *zeta;    // [K,D,h]
    double *x;             // [h, K*D] column-major (nullable when D == 0)
    double *ess_out;       // [K] nullable
    double *w_out;         // [K,P] nullable
    int32_t *comp_out;     // [K,D] nullable
};
cudaError_t launch_draw(const DrawArgs &a, cudaStream_t stream);

// HMC on the unconstrained hyperparameters with the leapfrog integrator on the device.
// Chains: K scenarios x P particles; slot vectors are [K, total] (total = theta_off[P]), per-chain scalars [K, P].
enum HmcStage : int { HMC_INIT = 0, HMC_INIT_DONE = 1, HMC_BEGIN = 2, HMC_LEAP_PRE = 3, HMC_LEAP_POST = 4, HMC_ACCEPT = 5 };
struct HmcArgs {
    int64_t K, P, total;
    int L;
    double eps;
    const int64_t *theta_off;                 // [P+1] device
    const int32_t *slot_kind;                 // [total]: 0 exp(a + b z), 2 2*logistic(a + b z), 3 z, 4 Phi(z), 5 constant a
    const double *slot_a, *slot_b;            // [total]
    int32_t noise_kind; double noise_a, noise_b;
    const double *momenta, *noise_momenta, *log_u;      // [n_steps, K, total], [n_steps, K, P], [n_steps, K, P]
    double *Z, *NZ;                           // current state
    double *Zq, *NZq, *mom, *mnz;             // trajectory
    double *gZ, *gNZ, *gq, *gnq;              // z-space gradients of the log posterior (current, trajectory)
    double *lp;                               // current log posterior
    double *theta, *noise;                    // constrained parameters the likelihood kernels read
    double *logml_q, *grad_theta, *grad_noise; int32_t *info_q;     // what they write
    double *logml_cur; int32_t *n_accept, *info_cur;                // outputs
    int32_t *iter;                            // iteration counter (device), read by every stage
};
cudaError_t launch_hmc_stage(const HmcArgs &h, int stage, int num_sms, cudaStream_t stream);

// Forecast summary: elementwise inverse transformation (x [h,N] column-major -> out same layout and/or
// rows [h][N]) and per-row type-7 quantiles by radix select.
cudaError_t launch_inverse_transform(int kind, double lam, double offset, double max_value, int64_t h, int64_t N,
                                     const double *x, double *out, double *rows, int num_sms, cudaStream_t stream);
cudaError_t launch_row_quantiles(int64_t h, int64_t N, const double *rows, int64_t nq, const double *probs, double *q,
                                 cudaStream_t stream);

}  // namespace nagp
