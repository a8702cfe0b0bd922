// Entry points lmc_nuts_sample / lmc_hmc_sample / lmc_workspace_bytes: argument validation and dispatch to the
// per-(target, kind) instantiations of the sampler kernel (lmc_sampler.cuh, lmc_inst_*.cu).
#include "lmc_sampler.cuh"

namespace lmc {

int check_sampler_args(const lmc_sampler_args* a, int kind, bool check_target) {
  if (!a) return LMC_ERR_BADARG;
  if (a->abi_version != LMC_ABI_VERSION) return LMC_ERR_BADARG;
  if (a->n_chains < 0 || a->ndim < 1 || a->n_trans < 0) return LMC_ERR_BADARG;
  if (a->ld < a->ndim || (a->ld & 1)) return LMC_ERR_BADARG;
  if (!a->q || !a->var || !a->adapt || !a->trace || !a->stats || !a->status || !a->workspace) return LMC_ERR_BADARG;
  if (((uintptr_t)a->q | (uintptr_t)a->var | (uintptr_t)a->workspace) & 15) return LMC_ERR_BADARG;
  if (a->adapt_mass) {
    if (!a->mean_fg || !a->rawvar_fg || !a->mean_bg || !a->rawvar_bg) return LMC_ERR_BADARG;
    if (((uintptr_t)a->mean_fg | (uintptr_t)a->rawvar_fg | (uintptr_t)a->mean_bg | (uintptr_t)a->rawvar_bg) & 15)
      return LMC_ERR_BADARG;
  }
  if (a->rng.mode == LMC_RNG_TAPE) {
    if (!a->rng.normals || !a->rng.uniforms || a->rng.u_stride < 1) return LMC_ERR_BADARG;
  } else if (a->rng.mode == LMC_RNG_PHILOX) {
    if (!a->rng.seeds) return LMC_ERR_BADARG;
  } else {
    return LMC_ERR_BADARG;
  }
  if (a->trace_skip < 0 || a->trace_skip > a->n_trans) return LMC_ERR_BADARG;
  if (a->progress && a->progress_block < 1) return LMC_ERR_BADARG;
  if (kind == KIND_NUTS) {
    if (a->tune_chunk != 0 && a->tune_chunk != 2 && a->tune_chunk != 4 && a->tune_chunk != 8 && a->tune_chunk != 16)
      return LMC_ERR_BADARG;
    if (a->max_treedepth < 1 || a->max_treedepth > kMaxDepth) return LMC_ERR_UNSUPPORTED;
    if (a->early_max_treedepth < 0 || a->early_max_treedepth > kMaxDepth) return LMC_ERR_UNSUPPORTED;
  } else {
    if (a->max_steps < 1) return LMC_ERR_BADARG;
  }
  if (!check_target) return LMC_OK;  // a user target compiled at run time (lmc_user.cu): args.target is ignored
  if (a->target.kind == LMC_TARGET_DIAG_GAUSSIAN) {
    if (!a->target.tau || ((uintptr_t)a->target.tau & 15)) return LMC_ERR_BADARG;
  } else if (a->target.kind != LMC_TARGET_FUNNEL) {
    return LMC_ERR_UNSUPPORTED;
  }
  return LMC_OK;
}

// one translation unit per (target, kind): lmc_inst_*.cu
int run_gauss_nuts(const lmc_sampler_args& a, const DiagGaussian& t);
int run_gauss_hmc(const lmc_sampler_args& a, const DiagGaussian& t);
int run_funnel_nuts(const lmc_sampler_args& a, const Funnel& t);
int run_funnel_hmc(const lmc_sampler_args& a, const Funnel& t);
// lean NUTS kernel (lmc_sampler_lean.cuh): forced with tune_group < 0 (threads per chain = -tune_group); the default
// for 513..1024 dimensions, where 128 threads x 4 pairs with three vectors in registers fit four chains per SM
// instead of three (measured +8% at 1024 chains x 1000 dimensions)
int run_gauss_nuts_lean(const lmc_sampler_args& a, const DiagGaussian& t, int group);
int run_funnel_nuts_lean(const lmc_sampler_args& a, const Funnel& t, int group);
bool pick_lean_shape(int ndim, int group, int* G, int* NP);
// chunked warp-per-chain NUTS kernel (lmc_sampler_warp.cuh): the default up to 256 dimensions (tune_group 0), forced
// with tune_group == 1
int run_gauss_nuts_warp(const lmc_sampler_args& a, const DiagGaussian& t);
int run_funnel_nuts_warp(const lmc_sampler_args& a, const Funnel& t);
// chunked CTA-per-chain NUTS kernel (lmc_sampler_cta.cuh): 128 threads per chain, up to 1024 dimensions; forced with
// tune_group == 2 (203 / 204: with 3 / 4 resident CTAs per SM)
int run_gauss_nuts_cta(const lmc_sampler_args& a, const DiagGaussian& t);
int run_funnel_nuts_cta(const lmc_sampler_args& a, const Funnel& t);
constexpr int kCtaRingMax = 4;  // largest position ring (leaves per chunk) of that kernel
static bool use_cta_kernel(int kind, int ndim, int tune_group) {
  return kind == KIND_NUTS && (ndim + 1) / 2 <= 512 && (tune_group == 2 || tune_group == 203 || tune_group == 204);
}
static bool use_warp_kernel(int kind, int ndim, int tune_group) {
  return kind == KIND_NUTS && (ndim + 1) / 2 <= 128 && (tune_group == 0 || tune_group == 1);
}
static int lean_group(const lmc_sampler_args* a, int kind) {
  if (kind != KIND_NUTS) return 0;
  if (a->tune_group < 0) return -a->tune_group;
  const int pairs = (a->ndim + 1) / 2;
  return (a->tune_group == 0 && pairs > 256 && pairs <= 512) ? 128 : 0;
}

static int sample_entry(const lmc_sampler_args* a, int kind) {
  const int rc = check_sampler_args(a, kind, true);
  if (rc != LMC_OK) return rc;
  if (a->n_chains == 0 || a->n_trans == 0) return LMC_OK;
  if (a->target.kind == LMC_TARGET_DIAG_GAUSSIAN) {
    DiagGaussian t{reinterpret_cast<const double2*>(a->target.tau)};
    if (use_cta_kernel(kind, a->ndim, a->tune_group)) return run_gauss_nuts_cta(*a, t);
    if (use_warp_kernel(kind, a->ndim, a->tune_group)) return run_gauss_nuts_warp(*a, t);
    if (lean_group(a, kind)) return run_gauss_nuts_lean(*a, t, lean_group(a, kind));
    return kind == KIND_NUTS ? run_gauss_nuts(*a, t) : run_gauss_hmc(*a, t);
  }
  Funnel t{1.0 / (a->target.v_scale * a->target.v_scale), 0.5 * (double)(a->ndim - 1)};
  if (use_cta_kernel(kind, a->ndim, a->tune_group)) return run_funnel_nuts_cta(*a, t);
  if (use_warp_kernel(kind, a->ndim, a->tune_group)) return run_funnel_nuts_warp(*a, t);
  if (lean_group(a, kind)) return run_funnel_nuts_lean(*a, t, lean_group(a, kind));
  return kind == KIND_NUTS ? run_funnel_nuts(*a, t) : run_funnel_hmc(*a, t);
}

}  // namespace lmc

extern "C" int64_t lmc_workspace_bytes(int32_t kind, int32_t n_chains, int32_t ndim, int32_t max_treedepth,
                                       int32_t tune_group) {
  if (n_chains < 0) return LMC_ERR_BADARG;
  if (kind == lmc::KIND_HMC) return (int64_t)lmc::sched_bytes(n_chains);
  if (kind != lmc::KIND_NUTS || max_treedepth < 1 || max_treedepth > lmc::kMaxDepth) return LMC_ERR_UNSUPPORTED;
  lmc::Shape s;
  if (lmc::use_cta_kernel(kind, ndim, tune_group)) {
    // one CTA of 128 threads per slot, at most 16 per SM; scratch = tree stack + trajectory + the position ring
    int dev = 0, n_sm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    (void)cudaGetLastError();
    const int np = (ndim + 1) / 2 <= 256 ? 2 : 4;
    long long slots = (long long)n_sm * 16;
    if (slots > n_chains) slots = n_chains < 1 ? 1 : n_chains;
    return (long long)lmc::sched_bytes(n_chains) + slots * (lmc::ws_vecs_nuts(max_treedepth) + lmc::kCtaRingMax) *
                                                       (long long)(128 * np) * (long long)sizeof(double2);
  }
  if (lmc::use_warp_kernel(kind, ndim, tune_group)) {
    // one warp per slot, at most 32 one-warp CTAs per SM; scratch = tree stack + trajectory + the largest position ring
    int dev = 0, n_sm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    (void)cudaGetLastError();
    const int np = (ndim + 1) / 2 <= 32 ? 1 : (ndim + 1) / 2 <= 64 ? 2 : 4;
    long long slots = (long long)n_sm * 32;
    if (slots > n_chains) slots = n_chains < 1 ? 1 : n_chains;
    return (long long)lmc::sched_bytes(n_chains) +
           slots * (lmc::ws_vecs_nuts(max_treedepth) + 16) * (long long)(32 * np) * (long long)sizeof(double2);
  }
  if (tune_group < 0) {
    if (!lmc::pick_lean_shape(ndim, -tune_group, &s.G, &s.NP)) return LMC_ERR_UNSUPPORTED;
  } else if (!lmc::pick_shape(ndim, tune_group, &s)) {
    return LMC_ERR_UNSUPPORTED;
  }
  // resident groups are bounded by 2048 threads per SM and by the number of chains
  int dev = 0, n_sm = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  (void)cudaGetLastError();
  const int cpb = s.G >= 64 ? 1 : 128 / s.G;
  long long slots = (long long)n_sm * (2048 / s.G);
  const long long by_chains = (((long long)n_chains + cpb - 1) / cpb) * cpb;
  if (slots > by_chains) slots = by_chains;
  if (slots < cpb) slots = cpb;
  return (long long)lmc::sched_bytes(n_chains) +
         slots * lmc::ws_vecs_nuts(max_treedepth) * (long long)(s.G * s.NP) * (long long)sizeof(double2);
}

extern "C" int lmc_nuts_sample(const lmc_sampler_args* args) { return lmc::sample_entry(args, lmc::KIND_NUTS); }
extern "C" int lmc_hmc_sample(const lmc_sampler_args* args) { return lmc::sample_entry(args, lmc::KIND_HMC); }
