// Host-side handle and user-model registry (the run-time analogue of the reference's
// link-time override of ssfunction/priorfun/checkbounds, external_inc.h:4-28).
#pragma once
#include <cuda_runtime.h>

#include <deque>
#include <string>
#include <vector>

#include "common.cuh"
#include "mcmcb200.h"

struct mcmcb_handle_s;

namespace mcmcb {

struct ModelEntry {
  const char* name;
  int kernel;  // 1 = small-npar register kernel, 2 = large-npar warp kernel
  int npar;    // compile-time npar, 0 = runtime
  int ny;
  int (*alloc)(mcmcb_handle_s*);
  int (*init)(mcmcb_handle_s*);
  int (*step)(mcmcb_handle_s*, int nsteps);
  int (*fetch)(mcmcb_handle_s*, const char* what, void* out, size_t out_bytes);
  int (*fetch_chain)(mcmcb_handle_s*, long long chain, int ld, double* chain_out, double* sschain_out,
                     double* s2chain_out, int* nrows);
  // pooled adaptation (pool.cuh): phase 1/2 = local moment sums into h->d_pool, 3 = apply the pooled factor
  int (*pool)(mcmcb_handle_s*, int phase);
};

std::vector<ModelEntry>& registry();
int register_model(const ModelEntry& e);

struct DumpSlot {
  double* dev = nullptr;
  double* host = nullptr;
  cudaEvent_t ev = nullptr, ready = nullptr;
  int state = 0;  // 0 free, 1 in flight / ready
  int step = 0;
};

}  // namespace mcmcb

struct mcmcb_handle_s {
  mcmcb_config cfg{};
  mcmcb::DevCfg dc{};
  int dodr = 0, doscam = 0, usesvd = 0;
  const mcmcb::ModelEntry* model = nullptr;
  int npar = 0, nycol = 0, L = 1;
  int num_sms = 0, occ = 1, blocks = 0;
  size_t max_smem = 0, smem = 0;
  bool attr_set = false, smem_blob = false, initial_set = false;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  long long pitch = 0, launches = 0, steps_done = 0;
  int nf = 0, inf = 0;
  double* d_st = nullptr;
  int* d_ist = nullptr;
  double *d_par0 = nullptr, *d_cmat0 = nullptr, *d_sigma2 = nullptr, *d_blob = nullptr, *d_prior = nullptr,
         *d_inj = nullptr;
  int* d_nobs = nullptr;
  size_t blob_n = 0, blob_bytes = 0, inj_per_chain = 0;
  int store_chains = 0;
  double *d_store_rows = nullptr, *d_store_cnt = nullptr, *d_store_s2 = nullptr;
  double* d_hist = nullptr;  // AP window ring of the register kernel (adapthist > 1)
  int hist_rows = 0;
  bool er_exit = false;  // method 'er': run the early-exit kernel (MCMCB_ER_EXIT=1)
  int k1_threads = 512;       // threads per CTA of the register kernel (<= MCMCB_K1_THREADS; MCMCB_K1_BLOCK overrides)
  bool k1_threads_fixed = false;  // MCMCB_K1_BLOCK given: no per-launch choice
  int k1_threads_used = 512;  // what the last launch used
  bool k1_supertile = true;   // bulk of a large population as per-warp super-tiles (MCMCB_K1_SUPERTILE=0: off)
  bool k1_exp_direct = true;  // stage the direct exp table in the shared memory left over (MCMCB_EXP_DIRECT=0: off)
  int k1_batch = 1;  // chains per thread of the register kernel (thread-per-chain mapping only)
  unsigned* d_tile = nullptr;
  // large-npar kernel (K2): per-chain vectors [chain][dp] and matrices [chain][d*d]
  double *d_theta = nullptr, *d_mean = nullptr, *d_Rm = nullptr, *d_cmat = nullptr, *d_rowbuf = nullptr,
         *d_scratch = nullptr, *d_cmat0_full = nullptr, *d_qstd = nullptr, *d_coef = nullptr, *d_Rp = nullptr;
  double *d_gcm = nullptr, *d_gmean = nullptr, *d_gw = nullptr;  // greedy burn-in accumulators (K2)
  int dp = 0, rowcap = 0, factor_mode = 0;
  long long r_stride = 0, q_stride = 0;
  bool r_resident = false;
  bool k4 = false;  // the thread-per-chain RAM kernel (k4_ram.cuh) ran the last launch
  int k5s_lanes = 0;  // lanes per chain of the last launch when it was k5s_scam_step_kernel, else 0
  int k2_warps = 8;
  int k2_group_threads = 0;  // > 0: the group-of-warps-per-chain kernel (k2g_group.cuh) runs, this many threads per chain
  long long k2_i = 1;  // simuind shared by all chains of the handle
  // pooled adaptation + diagnostics (pool.cuh, diag.cuh)
  mcmcb_allreduce_fn ar_fn = nullptr;
  void* ar_user = nullptr;
  double *d_pool = nullptr, *d_pool_partial = nullptr, *d_Rpool = nullptr;
  int* d_fail = nullptr;
  long long pool_ticks = 0;
  double *d_diag = nullptr, *d_diag_buf = nullptr, *d_diag_partial = nullptr;
  long long diag_n = 0;
  int diag_K = 0;
  void* d_fetch = nullptr;  // device staging of mcmcb_fetch (chain-major transposition happens on the device)
  size_t fetch_bytes = 0;
  std::vector<double> h_tmp;
  std::vector<mcmcb::DumpSlot> dump_slots;
  std::deque<int> dump_fifo;
  int dump_head = 0;
  long long dumps_dropped = 0;
  std::string err;
  // group handle (cfg.ngpus > 1): per-device handles and their NCCL communicators; no device state of its own
  std::vector<mcmcb_handle_s*> kids;
  std::vector<void*> nccl_comms;
  long long nccl_calls = 0;
};
