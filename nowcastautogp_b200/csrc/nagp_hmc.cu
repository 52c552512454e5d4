// nagp_hmc.cu — Hamiltonian Monte Carlo on the unconstrained hyperparameters, integrator on the device
// ("batched d logML / d theta + leapfrog on device").
//
// What it replaces: AutoGP.mcmc_parameters!(model, n_hmc) (/root/reference/src/forecasting.jl:148 and :65, and the
// n_hmc steps inside fit_smc!, /root/reference/src/make_and_fit_model.jl:91) runs Gen's HMC per particle on the
// CPU. Here every (scenario, particle) chain advances together and nothing returns to the host inside a chain:
// per leapfrog stage  [half kick + drift + z -> theta]  ->  tile kernel + gradient kernel  ->  [chain rule + half kick],
// and one accept/reject kernel per iteration. The momenta and the log-uniforms are supplied by the caller (like
// the normals of the forecast draws), so a host implementation fed the same numbers walks the same trajectory.
//
// Target: log p(y | theta(z)) + log N(z; 0, I); z -> theta per slot: 0 exp(a + b z) (log-normal prior),
// 2 2*logistic(a + b z), 3 z, 4 Phi(z), 5 constant a (not sampled). z is clipped to [-60, 60] inside the map so a
// diverging trajectory cannot overflow; it is rejected anyway. These small kernels are elementwise / one warp per
// chain: latency-bound glue between the two heavy kernels, which is why the whole iteration is one CUDA graph.
#include <algorithm>

#include "nagp_kernels.cuh"

namespace nagp {

namespace {

__device__ __forceinline__ double slot_theta(int kind, double a, double b, double z)
{
    z = fmin(fmax(z, -60.0), 60.0);
    switch (kind) {
    case 0: return exp(a + b * z);
    case 2: return 2.0 * (1.0 / (1.0 + exp(-(a + b * z))));
    case 3: return z;
    case 4: return 0.5 * (1.0 + erf(z / 1.4142135623730951));
    default: return a;
    }
}

__device__ __forceinline__ double slot_dtheta_dz(int kind, double b, double z, double theta)
{
    switch (kind) {
    case 0: return b * theta;
    case 2: return b * theta * (1.0 - 0.5 * theta);
    case 3: return 1.0;
    case 4: return exp(-0.5 * z * z) / 2.5066282746310002;
    default: return 0.0;
    }
}

__device__ __forceinline__ double finite_or_zero(double v) { return isfinite(v) ? v : 0.0; }

// elementwise over the K*total slots and the K*P noise parameters
__global__ void __launch_bounds__(256) hmc_elementwise_kernel(const HmcArgs h, const int stage)
{
    const int64_t nslot = h.K * h.total, nchain = h.K * h.P;
    const int64_t it = (stage == HMC_INIT) ? 0 : (int64_t)*h.iter;
    const bool learn_noise = h.noise_kind != 5;
    const double he = 0.5 * h.eps;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nslot + nchain;
         e += (int64_t)gridDim.x * blockDim.x) {
        if (e < nslot) {
            const int64_t j = e % h.total;
            const int kind = h.slot_kind[j];
            const double a = h.slot_a[j], b = h.slot_b[j];
            if (stage == HMC_INIT) {
                h.Zq[e] = h.Z[e];
                h.theta[e] = slot_theta(kind, a, b, h.Z[e]);
            } else if (stage == HMC_BEGIN) {
                h.mom[e] = h.momenta[it * nslot + e];
                h.Zq[e] = h.Z[e];
                h.gq[e] = h.gZ[e];
            } else if (stage == HMC_LEAP_PRE) {
                const double m = h.mom[e] + he * h.gq[e];
                const double zq = h.Zq[e] + h.eps * m;
                h.mom[e] = m;
                h.Zq[e] = zq;
                h.theta[e] = slot_theta(kind, a, b, zq);
            } else {   // HMC_LEAP_POST / HMC_INIT_DONE: chain rule to z-space and the N(0,1) prior, then the half kick
                const double zq = h.Zq[e];
                const double gr = finite_or_zero(h.grad_theta[e]) * slot_dtheta_dz(kind, b, zq, h.theta[e]) - zq;
                const double gq = finite_or_zero(gr);
                h.gq[e] = gq;
                if (stage == HMC_LEAP_POST) h.mom[e] = h.mom[e] + he * gq;
                else h.gZ[e] = gq;
            }
        } else {
            const int64_t c = e - nslot;
            if (stage == HMC_INIT) {
                h.NZq[c] = h.NZ[c];
                h.noise[c] = slot_theta(h.noise_kind, h.noise_a, h.noise_b, h.NZ[c]);
            } else if (stage == HMC_BEGIN) {
                h.mnz[c] = learn_noise ? h.noise_momenta[it * nchain + c] : 0.0;
                h.NZq[c] = h.NZ[c];
                h.gnq[c] = h.gNZ[c];
            } else if (stage == HMC_LEAP_PRE) {
                if (learn_noise) {
                    const double m = h.mnz[c] + he * h.gnq[c];
                    const double zq = h.NZq[c] + h.eps * m;
                    h.mnz[c] = m;
                    h.NZq[c] = zq;
                    h.noise[c] = slot_theta(h.noise_kind, h.noise_a, h.noise_b, zq);
                }
            } else {
                double gq = 0.0;
                if (learn_noise) {
                    const double zq = h.NZq[c];
                    gq = finite_or_zero(finite_or_zero(h.grad_noise[c]) *
                                        slot_dtheta_dz(h.noise_kind, h.noise_b, zq, h.noise[c]) - zq);
                }
                h.gnq[c] = gq;
                if (stage == HMC_LEAP_POST) h.mnz[c] = h.mnz[c] + he * gq;
                else h.gNZ[c] = gq;
            }
        }
    }
    if (stage == HMC_INIT && blockIdx.x == 0 && threadIdx.x == 0) *h.iter = 0;
}

// one warp per chain: prior and kinetic energies over the chain's slots, then accept / reject (or, for
// HMC_INIT_DONE, adopt the initial evaluation as the current state)
__global__ void __launch_bounds__(256) hmc_chain_kernel(const HmcArgs h, const int stage)
{
    const int lane = threadIdx.x & 31;
    const int64_t nchain = h.K * h.P, nslot = h.K * h.total;
    const int64_t it = (stage == HMC_INIT_DONE) ? 0 : (int64_t)*h.iter;
    const bool learn_noise = h.noise_kind != 5;
    for (int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < nchain;
         c += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t s = c / h.P;
        const int p = (int)(c % h.P);
        const int64_t o = s * h.total + h.theta_off[p];
        const int ns = (int)(h.theta_off[p + 1] - h.theta_off[p]);
        double zz = 0.0, k1 = 0.0, k0 = 0.0;
        for (int j = lane; j < ns; j += 32) {
            const double zq = h.Zq[o + j];
            zz += zq * zq;
            if (stage == HMC_ACCEPT) {
                const double m1 = h.mom[o + j], m0 = h.momenta[it * nslot + o + j];
                k1 += m1 * m1; k0 += m0 * m0;
            }
        }
        for (int d = 16; d > 0; d >>= 1) {
            zz += __shfl_xor_sync(0xffffffffu, zz, d);
            k1 += __shfl_xor_sync(0xffffffffu, k1, d);
            k0 += __shfl_xor_sync(0xffffffffu, k0, d);
        }
        const double nzq = h.NZq[c];
        double prior = -0.5 * zz;
        if (learn_noise) prior -= 0.5 * nzq * nzq;
        const int inf = h.info_q[c];
        const double lm = h.logml_q[c];
        const double lp_new = inf == 0 ? lm + prior : -INFINITY;
        bool accept;
        if (stage == HMC_INIT_DONE) {
            accept = true;
        } else {
            double K1 = 0.5 * k1, K0 = 0.5 * k0;
            if (learn_noise) {
                const double m1 = h.mnz[c], m0 = h.noise_momenta[it * nchain + c];
                K1 += 0.5 * m1 * m1; K0 += 0.5 * m0 * m0;
            }
            const double h0 = -h.lp[c] + K0, h1 = -lp_new + K1;
            accept = isfinite(h1) && h.log_u[it * nchain + c] < h0 - h1;
        }
        if (accept) {
            for (int j = lane; j < ns; j += 32) { h.Z[o + j] = h.Zq[o + j]; h.gZ[o + j] = h.gq[o + j]; }
            if (lane == 0) {
                h.NZ[c] = nzq; h.gNZ[c] = h.gnq[c]; h.lp[c] = lp_new;
                h.logml_cur[c] = inf == 0 ? lm : nan("");
                h.info_cur[c] = inf;
                if (stage == HMC_INIT_DONE) h.n_accept[c] = 0; else h.n_accept[c] += 1;
            }
        }
    }
}

__global__ void hmc_next_kernel(int32_t *iter) { *iter += 1; }

}  // namespace

cudaError_t launch_hmc_stage(const HmcArgs &h, int stage, int num_sms, cudaStream_t stream)
{
    const int64_t work = h.K * h.total + h.K * h.P;
    const int grid_e = (int)std::min<int64_t>((work + 255) / 256, (int64_t)num_sms * 8);
    const int grid_c = (int)std::min<int64_t>((h.K * h.P + 7) / 8, (int64_t)num_sms * 8);
    if (stage == HMC_INIT || stage == HMC_BEGIN || stage == HMC_LEAP_PRE || stage == HMC_LEAP_POST) {
        hmc_elementwise_kernel<<<grid_e, 256, 0, stream>>>(h, stage);
    } else if (stage == HMC_INIT_DONE) {
        hmc_elementwise_kernel<<<grid_e, 256, 0, stream>>>(h, stage);      // gZ from the initial gradient
        hmc_chain_kernel<<<grid_c, 256, 0, stream>>>(h, stage);
    } else {
        hmc_chain_kernel<<<grid_c, 256, 0, stream>>>(h, stage);
        hmc_next_kernel<<<1, 1, 0, stream>>>(h.iter);
    }
    return cudaGetLastError();
}

}  // namespace nagp
