// host_solver_bench.cpp - the REPLICATED host share of a batch solve, measured without a GPU: K = 400 batch-shaped problem
// (n = 2400, half bandwidth 41), between chain + prior from the stand-in host factors + synthetic 6x6 pair blocks, through the
// product minimizer.  Build and run: scripts/build_host_bench.sh && /tmp/host_solver_bench
#include <chrono>
#include <cstdio>
#include <random>
#include "solver.h"
#include "../../include/glio_b200.h"   // relative to glio_b200/csrc (-I)
using namespace glio;
using clk = std::chrono::steady_clock;
int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 400, sr = 6, nt = 6, na = 7, n = K * nt, hb = (sr + 1) * nt - 1;
  std::mt19937 rng(3); std::normal_distribution<double> N01(0, 1);
  glio_host_factor_set* hf = glio_hf_create();
  std::vector<double> T(K * 7);
  for (int k = 0; k < K; ++k) { T[7*k] = 0.5 * k; T[7*k+1] = 0.1*std::sin(0.1*k); T[7*k+2] = 0; double a = 0.01 * k; T[7*k+3] = std::cos(a/2); T[7*k+4] = 0; T[7*k+5] = 0; T[7*k+6] = std::sin(a/2); }
  double sw[15]; for (int i = 0; i < 3; ++i) sw[i] = 10; for (int i = 3; i < 6; ++i) sw[i] = 30; for (int i = 6; i < 15; ++i) sw[i] = 0;
  glio_hf_add_prior(hf, 0, &T[0], &T[3], nullptr, sw);
  for (int i = 0; i + 1 < K; ++i) {
    // dp = R_i^T (t_{i+1}-t_i), dq = conj(q_i) q_{i+1}
    const double* qi = &T[7*i+3]; const double* qj = &T[7*i+10];
    double d[3] = {T[7*i+7]-T[7*i], T[7*i+8]-T[7*i+1], T[7*i+9]-T[7*i+2]};
    double a = 2*std::atan2(qi[3], qi[0]); double c = std::cos(-a), s = std::sin(-a);
    double dp[3] = {c*d[0]-s*d[1] + 0.005*N01(rng), s*d[0]+c*d[1] + 0.005*N01(rng), d[2] + 0.005*N01(rng)};
    double da = 2*std::atan2(qj[3], qj[0]) - a; double dq[4] = {std::cos(da/2), 0, 0, std::sin(da/2)};
    double dv[3] = {0,0,0};
    glio_hf_add_between(hf, i, i+1, dp, dq, dv, 0.1, sw);
  }
  // synthetic LiDAR pair blocks: residual r_p = A_p (dx_c - dx_o) in tangent-ish linearisation around truth: cost 0.5 w |t_c - t_o - (T_c - T_o)|^2
  std::vector<std::pair<int,int>> pairs;
  for (int c = 0; c < K; ++c) for (int o = c + 1; o <= std::min(K - 1, c + sr); ++o) { pairs.push_back({c, o}); pairs.push_back({o, c}); }
  std::vector<ParamBlock> blocks;
  for (int k = 0; k < K; ++k) { blocks.push_back(ParamBlock{na*k, 3, nt*k, 3, false, nullptr}); blocks.push_back(ParamBlock{na*k+3, 4, nt*k+3, 3, true, nullptr}); }
  SolverOptions so; so.max_num_iterations = 100; so.dogleg_type = 1; so.use_nonmonotonic_steps = true;
  double t_hf = 0, t_asm = 0; int nev = 0;
  std::vector<double> pz(K * 7);
  EvalFn eval = [&](const double* xa, bool wj, double* cost, BandMat* H, double* g) -> bool {
    ++nev;
    for (int k = 0; k < K; ++k) for (int i = 0; i < 7; ++i) pz[7*k+i] = xa[na*k+i];
    double ct = 0;
    auto t0 = clk::now();
    if (wj) { H->reset(n, hb); std::fill(g, g + n, 0.0); }
    const double w = 100.0;
    for (auto& p : pairs) {
      const int c = p.first, o = p.second;
      double r[3]; for (int i = 0; i < 3; ++i) r[i] = (pz[7*c+i] - pz[7*o+i]) - (T[7*c+i] - T[7*o+i]);
      ct += 0.5 * w * (r[0]*r[0] + r[1]*r[1] + r[2]*r[2]);
      if (wj) {
        double ob[36] = {0}; for (int i = 0; i < 3; ++i) ob[6*i+i] = -w;
        for (int i = 0; i < 3; ++i) { H->at(nt*c+i, nt*c+i) += w; H->at(nt*o+i, nt*o+i) += w; g[nt*c+i] += w * r[i]; g[nt*o+i] -= w * r[i]; }
        for (int pi = 0; pi < 6; ++pi) for (int q = 0; q < 6; ++q) H->add_sym(nt*c+pi, nt*o+q, ob[6*pi+q]);
      }
    }
    auto t1 = clk::now();
    if (glio_hf_evaluate_band(hf, K, pz.data(), nullptr, wj ? 1 : 0, wj ? H->a.data() : nullptr, wj ? H->hb : 0, g, &ct) != 0) return false;
    auto t2 = clk::now();
    t_asm += std::chrono::duration<double>(t1 - t0).count(); t_hf += std::chrono::duration<double>(t2 - t1).count();
    *cost = ct; return std::isfinite(ct);
  };
  for (int rep = 0; rep < 3; ++rep) {
    std::vector<double> x(K * na);
    std::mt19937 r2(5);
    for (int k = 0; k < K; ++k) { for (int i = 0; i < 7; ++i) x[na*k+i] = T[7*k+i]; for (int i = 0; i < 3; ++i) x[na*k+i] += 0.05 * N01(r2); }
    t_hf = t_asm = 0; nev = 0;
    TrustRegionDogleg solver(blocks, so);
    SolverSummary S; solver.solve(x.data(), eval, &S);
    printf("iters %zu evals %d lin %d | total %.3f ms eval %.3f (asm %.3f hf %.3f) linear %.3f other %.3f | %s cost %.4g -> %.4g\n", S.iterations.size(), nev, S.num_linear_solves,
           1e3*S.total_seconds, 1e3*S.eval_seconds, 1e3*t_asm, 1e3*t_hf, 1e3*S.linear_solver_seconds, 1e3*(S.total_seconds - S.eval_seconds - S.linear_solver_seconds), S.message.c_str(), S.initial_cost, S.final_cost);
  }
}
