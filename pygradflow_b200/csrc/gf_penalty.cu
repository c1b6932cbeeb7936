// Penalty strategies that look at the candidate iterate (pygradflow/penalty.py:115-255) for the batched driver.
//
//   gf_pareto_update : ParetoDecrease.update (penalty.py:136-168) -- never rejects; evaluated on the committed iterate
//                      (the products J'c and J'y are the ones the termination test needs anyway).
//   gf_filter_update : ObjectivePenaltyFilter / LagrangianPenaltyFilter (penalty.py:170-255) + the veto of
//                      solver.py:357-378: an instance whose accepted candidate is dominated by an entry of its filter is
//                      turned into a rejected step (phase <- GF_PHASE_REJECT: the iterate stays, lambda and the iteration
//                      counter move on) and the STRATEGY's rho is multiplied by 10, while the solver's rho only follows
//                      at the next accepted step -- exactly the reference's bookkeeping.
// One CTA per instance; reductions on warp shuffles in a fixed order.
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

constexpr int PEN_THREADS = 128;

__global__ void __launch_bounds__(PEN_THREADS) pareto_kernel(int n, int m, const double* __restrict__ grad,
                                                             const double* __restrict__ cons,
                                                             const double* __restrict__ jty,
                                                             const double* __restrict__ jtc,
                                                             const int32_t* __restrict__ phase,
                                                             const int32_t* __restrict__ status, double opt_tol,
                                                             double local_infeas_tol, double* __restrict__ rho,
                                                             GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int ph = phase[b];
    if (status[b] != 0 || (ph != GF_PHASE_ACCEPT_MID && ph != GF_PHASE_ACCEPT_FINAL)) return;
    __shared__ double red[32];
    const double* g = grad + (size_t)b * n;
    const double* c = cons + (size_t)b * m;
    const double* ty = jty + (size_t)b * n;
    const double* tc = jtc + (size_t)b * n;
    double cc = 0.0;
    for (int j = threadIdx.x; j < m; j += PEN_THREADS) cc = fma(c[j], c[j], cc);
    cc = block_sum(cc, red);
    const double viol = 0.5 * cc;
    if (viol <= opt_tol) return;  // already feasible
    double tcmax = 0.0, gtc = 0.0, gg = 0.0, tyg = 0.0, tcs = 0.0, tctc = 0.0;
    for (int i = threadIdx.x; i < n; i += PEN_THREADS) {
        const double gi = g[i], yi = ty[i], ci = tc[i];
        tcmax = fmax(tcmax, fabs(ci));
        gtc = fma(gi, ci, gtc);
        gg = fma(gi, gi, gg);
        tyg = fma(yi, gi, tyg);
        tcs = fma(ci, gi + yi, tcs);
        tctc = fma(ci, ci, tctc);
    }
    tcmax = block_max(tcmax, red);
    if (tcmax <= local_infeas_tol) return;  // locally infeasible: no bound
    gtc = block_sum(gtc, red);
    gg = block_sum(gg, red);
    tyg = block_sum(tyg, red);
    tcs = block_sum(tcs, red);
    tctc = block_sum(tctc, red);
    if (threadIdx.x == 0) {
        double obj_bound = INFINITY;
        if (fabs(gtc) > 1e-10) obj_bound = -(sqrt(gg) + tyg) / gtc;
        const double cons_bound = -tcs / sqrt(tctc);
        const double bound = fmin(obj_bound, cons_bound);
        const double r = rho[b];
        if (isfinite(bound)) rho[b] = fmax(fmin(r * 10.0, bound), r);  // the reference asserts finiteness
    }
}

// kind 0: ObjectiveFilter entry (obj, |c|_inf); kind 1: LagrangianFilter entry (|dL|^2 + |c|^2, |c|_2) with dL the
// augmented-Lagrangian gradient of the candidate at the STRATEGY's rho (the caller evaluates it).
__global__ void __launch_bounds__(PEN_THREADS) filter_kernel(int n, int m, int kind, int32_t* __restrict__ phase,
                                                             const int32_t* __restrict__ status,
                                                             const double* __restrict__ om, const double* __restrict__ cm,
                                                             const double* __restrict__ dm, const double* __restrict__ of,
                                                             const double* __restrict__ cf, const double* __restrict__ df,
                                                             double* __restrict__ rho, double* __restrict__ rho_pen,
                                                             double* __restrict__ filt, int32_t* __restrict__ nfilt,
                                                             int cap, int32_t* __restrict__ overflow, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int ph = phase[b];
    if (status[b] != 0 || (ph != GF_PHASE_ACCEPT_MID && ph != GF_PHASE_ACCEPT_FINAL)) return;
    __shared__ double red[32];
    __shared__ int dom;
    const bool mid = ph == GF_PHASE_ACCEPT_MID;
    const double* c = (mid ? cm : cf) + (size_t)b * m;
    double first, second;
    if (kind == 0) {
        double cmax = 0.0;
        for (int j = threadIdx.x; j < m; j += PEN_THREADS) cmax = fmax(cmax, fabs(c[j]));
        cmax = block_max(cmax, red);
        first = (mid ? om : of)[b];
        second = m > 0 ? cmax : 0.0;
    } else {
        const double* d = (mid ? dm : df) + (size_t)b * n;
        double dd = 0.0, cc = 0.0;
        for (int i = threadIdx.x; i < n; i += PEN_THREADS) dd = fma(d[i], d[i], dd);
        for (int j = threadIdx.x; j < m; j += PEN_THREADS) cc = fma(c[j], c[j], cc);
        dd = block_sum(dd, red);
        cc = block_sum(cc, red);
        first = dd + cc;
        second = sqrt(cc);
    }
    double* F = filt + (size_t)b * cap * 2;
    const int cnt = nfilt[b];
    if (threadIdx.x == 0) dom = 0;
    __syncthreads();
    int mine = 0;
    for (int e = threadIdx.x; e < cnt; e += PEN_THREADS) mine |= (F[2 * e] <= first && F[2 * e + 1] <= second) ? 1 : 0;
    if (mine) atomicOr(&dom, 1);
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (dom) {  // penalty.py:209-210 + solver.py:360: reject, the strategy's rho grows, the solver's does not
        rho_pen[b] *= 10.0;
        phase[b] = GF_PHASE_REJECT;
        return;
    }
    int k = 0;
    for (int e = 0; e < cnt; ++e) {  // drop the entries the new one dominates (order kept)
        const double e0 = F[2 * e], e1 = F[2 * e + 1];
        if (!(first <= e0 && second <= e1)) {
            F[2 * k] = e0;
            F[2 * k + 1] = e1;
            ++k;
        }
    }
    if (k < cap) {
        F[2 * k] = first;
        F[2 * k + 1] = second;
        ++k;
    } else {
        overflow[b] = 1;
    }
    nfilt[b] = k;
    rho[b] = rho_pen[b];  // solver.py:364-369
}

}  // namespace

extern "C" {

int gf_pareto_update(int B, int n, int m, const double* grad, const double* cons, const double* jty, const double* jtc,
                     const int32_t* phase, const int32_t* status, double opt_tol, double local_infeas_tol, double* rho,
                     const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B < 0 || n <= 0 || m < 0 || !grad || !phase || !status || !rho || nwork < 0) return GF_ERR_ARG;
    if (m > 0 && (!cons || !jty || !jtc)) return GF_ERR_ARG;
    if (nwork == 0 || B == 0 || m == 0) return GF_OK;  // no constraints: viol = 0 <= opt_tol, rho stays
    pareto_kernel<<<nwork, PEN_THREADS, 0, (cudaStream_t)stream>>>(n, m, grad, cons, jty, jtc, phase, status, opt_tol,
                                                                   local_infeas_tol, rho, GfWork{work, nwork_dev});
    return gf_launch_status();
}

int gf_filter_update(int B, int n, int m, int kind, int32_t* phase, const int32_t* status, const double* om,
                     const double* cm, const double* dm, const double* of, const double* cf, const double* df,
                     double* rho, double* rho_pen, double* filt, int32_t* nfilt, int cap, int32_t* overflow,
                     const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B < 0 || n <= 0 || m < 0 || (kind != 0 && kind != 1) || !phase || !status || !rho || !rho_pen || !filt || !nfilt ||
        cap < 1 || !overflow || nwork < 0)
        return GF_ERR_ARG;
    if (kind == 0 && (!om || !of)) return GF_ERR_ARG;
    if (kind == 1 && (!dm || !df)) return GF_ERR_ARG;
    if (m > 0 && (!cm || !cf)) return GF_ERR_ARG;
    if (nwork == 0 || B == 0) return GF_OK;
    filter_kernel<<<nwork, PEN_THREADS, 0, (cudaStream_t)stream>>>(n, m, kind, phase, status, om, cm, dm, of, cf, df, rho,
                                                                   rho_pen, filt, nfilt, cap, overflow,
                                                                   GfWork{work, nwork_dev});
    return gf_launch_status();
}

}  // extern "C"
