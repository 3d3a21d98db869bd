// shim_solve.h — implementation of ceres::Solve / Problem::Evaluate for the glio_b200 Ceres-API shim (see ceres.h).
#ifndef GLIO_SHIM_CERES_SHIM_SOLVE_H_
#define GLIO_SHIM_CERES_SHIM_SOLVE_H_

#include <chrono>

namespace ceres {
namespace shim_internal {

struct ActiveBlock { double* user; int size, local, amb_off, tan_off; LocalParameterization* param; };

struct Program {
  std::vector<ActiveBlock> active;
  std::map<double*, int> index;           // user pointer -> active index
  int n_amb = 0, n_tan = 0;
};

// Ceres' Corrector (ceres.tgz::internal/ceres/corrector.cc) for a residual block with nr rows
struct Corr {
  double sqrt_rho1 = 1, residual_scaling = 1, alpha_sq_norm = 0;
  Corr(double sq_norm, const double rho[3]) {
    sqrt_rho1 = std::sqrt(rho[1]);
    if (sq_norm == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; return; }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  void jac(int nr, int nc, const double* r, double* J) const {
    if (alpha_sq_norm == 0.0) { for (int i = 0; i < nr * nc; ++i) J[i] *= sqrt_rho1; return; }
    for (int c = 0; c < nc; ++c) {
      double rtj = 0; for (int k = 0; k < nr; ++k) rtj += J[k * nc + c] * r[k];
      for (int k = 0; k < nr; ++k) J[k * nc + c] = sqrt_rho1 * (J[k * nc + c] - alpha_sq_norm * r[k] * rtj);
    }
  }
};

inline bool is_f32(double v) { return (double)(float)v == v; }

}  // namespace shim_internal

inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  using namespace shim_internal;
  const auto t_start = std::chrono::steady_clock::now();
  Solver::Summary& S = *summary;
  S = Solver::Summary();
  // ---- reduced program (Program::RemoveFixedBlocks, ceres.tgz::internal/ceres/program.cc:296-366) ----
  std::set<double*> used;
  for (auto& rb : problem->residual_blocks_) for (double* p : rb->params) used.insert(p);
  Program P;
  for (double* u : problem->order_) {
    const Problem::Block& b = problem->blocks_.at(u);
    if (b.constant || !used.count(u)) continue;
    ActiveBlock a; a.user = u; a.size = b.size; a.param = b.param; a.local = b.param ? b.param->LocalSize() : b.size;
    a.amb_off = P.n_amb; a.tan_off = P.n_tan; P.n_amb += a.size; P.n_tan += a.local;
    P.index[u] = (int)P.active.size(); P.active.push_back(a);
  }
  S.num_parameter_blocks = (int)P.active.size(); S.num_parameters = P.n_amb; S.num_effective_parameters = P.n_tan;
  S.num_residual_blocks = problem->NumResidualBlocks(); S.num_residuals = problem->NumResiduals();
  const int n = P.n_tan;
  if (n == 0) { S.termination_type = CONVERGENCE; S.message = "no free parameters"; S.initial_cost = S.final_cost = 0; return; }

  // ---- split residual blocks: device-evaluable unary plane factors vs host ----
  struct Slot { double* t; double* q; std::vector<float> cp, nsd, w; };
  std::vector<Slot> slots;
  std::map<std::pair<double*, double*>, int> slot_of;
  std::vector<const Problem::Residual*> host_blocks;
  struct BPair { int kc, ko; std::vector<float> cp; std::vector<double> nc; std::vector<float> w; };
  std::vector<BPair> bpairs;                                     // in creation order == glio_batch_pair_list order
  std::map<std::pair<int, int>, int> bpair_of;
  std::vector<std::pair<double*, double*>> bkfs;                 // keyframes (t, q blocks) touched by binary factors
  std::map<std::pair<double*, double*>, int> bkf_of;
  glio_ctx* ctx = problem->ctx_;
  glio_params gp;
  if (ctx && glio_get_params(ctx, &gp) != GLIO_OK) ctx = nullptr;
  for (auto& rbp : problem->residual_blocks_) {
    const Problem::Residual* rb = rbp.get();
    glio::FactorDesc d;
    bool dev = false;
    const bool described = ctx && rb->cost->GlioDescribe(&d);
    if (!described) d.kind = glio::FACTOR_NONE;
    if (described && d.kind == glio::FACTOR_PLANE_UNARY && rb->params.size() == 2) {
      double* t = rb->params[0]; double* q = rb->params[1];
      const HuberLoss* hl = dynamic_cast<const HuberLoss*>(rb->loss);
      const bool loss_ok = (rb->loss == nullptr && gp.huber_delta <= 0) || (hl && hl->a() == gp.huber_delta);
      const bool blocks_ok = P.index.count(t) && P.index.count(q) && problem->blocks_.at(t).size == 3 && problem->blocks_.at(q).size == 4 &&
                             dynamic_cast<QuaternionParameterization*>(problem->blocks_.at(q).param) != nullptr;
      bool ext_ok = true;
      for (int k = 0; k < 4; ++k) ext_ok = ext_ok && d.q_lb[k] == gp.q_lb[k];
      for (int k = 0; k < 3; ++k) ext_ok = ext_ok && d.t_lb[k] == gp.t_lb[k];
      // unit_score contexts evaluate every match with score = lidar_const (front-end factor): only blocks with exactly that score qualify
      const double w = gp.unit_score ? 1.0 : d.score / gp.lidar_const;
      const bool repr = (gp.unit_score ? d.score == gp.lidar_const : (is_f32(w) && gp.lidar_const * (double)(float)w == d.score)) && is_f32(d.cp[0]) && is_f32(d.cp[1]) && is_f32(d.cp[2]) &&
                        is_f32(d.n[0]) && is_f32(d.n[1]) && is_f32(d.n[2]) && is_f32(d.d);
      if (loss_ok && blocks_ok && ext_ok && repr) {
        auto key = std::make_pair(t, q);
        auto it = slot_of.find(key);
        int si;
        if (it == slot_of.end()) { si = (int)slots.size(); slot_of[key] = si; slots.push_back(Slot{t, q, {}, {}, {}}); } else si = it->second;
        Slot& sl = slots[si];
        for (int k = 0; k < 3; ++k) sl.cp.push_back((float)d.cp[k]);
        for (int k = 0; k < 3; ++k) sl.nsd.push_back((float)d.n[k]);
        sl.nsd.push_back((float)d.d); sl.w.push_back((float)w);
        dev = true;
      }
    }
    // BinaryLidarPlaneNormFactor (LidarKeyframeFactor.h:124-164): four blocks (t_cur, q_cur, t_oth, q_oth), no loss
    // (Estimator.cpp:2768); grouped by keyframe pair and evaluated by K2b
    if (!dev && ctx && d.kind == glio::FACTOR_PLANE_BINARY && rb->params.size() == 4 && rb->loss == nullptr && gp.batch_score != 0.0) {
      double* tc = rb->params[0]; double* qc = rb->params[1]; double* to = rb->params[2]; double* qo = rb->params[3];
      auto blk_ok = [&](double* t, double* q) {
        return P.index.count(t) && P.index.count(q) && problem->blocks_.at(t).size == 3 && problem->blocks_.at(q).size == 4 &&
               dynamic_cast<QuaternionParameterization*>(problem->blocks_.at(q).param) != nullptr;
      };
      const double w = d.score / gp.batch_score;
      const bool repr = is_f32(w) && gp.batch_score * (double)(float)w == d.score && is_f32(d.cp[0]) && is_f32(d.cp[1]) && is_f32(d.cp[2]);
      if (blk_ok(tc, qc) && blk_ok(to, qo) && !(tc == to && qc == qo) && repr) {
        auto kf_of = [&](double* t, double* q) {
          auto key = std::make_pair(t, q);
          auto it = bkf_of.find(key);
          if (it != bkf_of.end()) return it->second;
          const int k = (int)bkfs.size(); bkf_of[key] = k; bkfs.push_back(key); return k;
        };
        const int kc = kf_of(tc, qc), ko = kf_of(to, qo);
        auto pkey = std::make_pair(kc, ko);
        auto pit = bpair_of.find(pkey);
        int pi;
        if (pit == bpair_of.end()) { pi = (int)bpairs.size(); bpair_of[pkey] = pi; bpairs.push_back(BPair{kc, ko, {}, {}, {}}); } else pi = pit->second;
        BPair& bp = bpairs[pi];
        for (int k = 0; k < 3; ++k) bp.cp.push_back((float)d.cp[k]);
        for (int k = 0; k < 6; ++k) bp.nc.push_back(d.nc[k]);
        bp.w.push_back((float)w);
        dev = true;
      }
    }
    if (!dev) host_blocks.push_back(rb);
  }
  const int W = (int)slots.size();
  const int KB = (int)bkfs.size(), PB = (int)bpairs.size();
  if (PB > 0) {
    if (glio_batch_clear(ctx) != GLIO_OK) { S.termination_type = FAILURE; S.message = std::string("glio_batch_clear: ") + glio_last_error(ctx); return; }
    for (int p = 0; p < PB; ++p) {
      if (glio_batch_set_pair_matches(ctx, bpairs[p].kc, bpairs[p].ko, bpairs[p].cp.data(), bpairs[p].nc.data(), bpairs[p].w.data(), (int64_t)bpairs[p].w.size()) != GLIO_OK) {
        S.termination_type = FAILURE; S.message = std::string("glio_batch_set_pair_matches: ") + glio_last_error(ctx); return;
      }
      S.num_device_residual_blocks += (int)bpairs[p].w.size();
    }
  }
  for (int k = 0; k < W; ++k) {
    if (glio_set_matches(ctx, k, slots[k].cp.data(), slots[k].nsd.data(), slots[k].w.data(), (int64_t)slots[k].w.size()) != GLIO_OK) {
      S.termination_type = FAILURE; S.message = std::string("glio_set_matches: ") + glio_last_error(ctx); return;
    }
    S.num_device_residual_blocks += (int)slots[k].w.size();
  }

  // ---- band of J^T J from the block structure ----
  int hb = 0;
  for (auto& rbp : problem->residual_blocks_) {
    int lo = n, hi = -1;
    for (double* p : rbp->params) { auto it = P.index.find(p); if (it == P.index.end()) continue; const ActiveBlock& a = P.active[it->second]; lo = std::min(lo, a.tan_off); hi = std::max(hi, a.tan_off + a.local - 1); }
    if (hi >= 0) hb = std::max(hb, hi - lo);
  }

  // ---- solver setup ----
  std::vector<glio::ParamBlock> blocks;
  for (const ActiveBlock& a : P.active) {
    glio::ParamBlock b{a.amb_off, a.size, a.tan_off, a.local, false, nullptr};
    if (a.param) { LocalParameterization* lp = a.param; b.plus_fn = [lp](const double* x, const double* d, double* o) { lp->Plus(x, d, o); }; }
    blocks.push_back(b);
  }
  glio::SolverOptions so;
  so.max_num_iterations = options.max_num_iterations;
  so.max_solver_time_in_seconds = options.max_solver_time_in_seconds;
  so.dogleg_type = options.dogleg_type == SUBSPACE_DOGLEG ? 1 : 0;
  so.use_nonmonotonic_steps = options.use_nonmonotonic_steps;
  so.max_consecutive_nonmonotonic_steps = options.max_consecutive_nonmonotonic_steps;
  so.initial_trust_region_radius = options.initial_trust_region_radius; so.max_trust_region_radius = options.max_trust_region_radius;
  so.min_trust_region_radius = options.min_trust_region_radius; so.min_relative_decrease = options.min_relative_decrease;
  so.min_lm_diagonal = options.min_lm_diagonal; so.max_lm_diagonal = options.max_lm_diagonal;
  so.max_num_consecutive_invalid_steps = options.max_num_consecutive_invalid_steps; so.jacobi_scaling = options.jacobi_scaling;
  so.function_tolerance = options.function_tolerance; so.gradient_tolerance = options.gradient_tolerance; so.parameter_tolerance = options.parameter_tolerance;
  // LEVENBERG_MARQUARDT is Ceres' default and what the front end's scan matcher runs (LidarOdometry.cpp:521-530); DOGLEG is
  // what the Estimator selects (Estimator.cpp:2427, :3278).  Both are implemented by the host minimizer.
  so.trust_region_strategy = options.trust_region_strategy_type == LEVENBERG_MARQUARDT ? 1 : 0;
  glio::TrustRegionDogleg solver(blocks, so);

  std::vector<double> x(P.n_amb);
  for (const ActiveBlock& a : P.active) std::memcpy(&x[a.amb_off], a.user, sizeof(double) * a.size);
  std::vector<double> poses((size_t)std::max(W, 1) * 7), Hd((size_t)std::max(W, 1) * 36), gd((size_t)std::max(W, 1) * 6), cd(std::max(W, 1));
  std::vector<double> bposes((size_t)std::max(KB, 1) * 7), bHd((size_t)std::max(KB, 1) * 36), bHo((size_t)std::max(PB, 1) * 36), bg((size_t)std::max(KB, 1) * 6);
  double fixed_cost = 0; bool fixed_done = false;

  glio::EvalFn eval = [&](const double* xa, bool want_jac, double* cost, glio::BandMat* H, double* g) -> bool {
    double ct = 0;
    if (want_jac) { H->reset(n, hb); std::fill(g, g + n, 0.0); }
    // device blocks
    if (W > 0) {
      for (int k = 0; k < W; ++k) {
        const ActiveBlock& at = P.active[P.index[slots[k].t]]; const ActiveBlock& aq = P.active[P.index[slots[k].q]];
        for (int i = 0; i < 3; ++i) poses[7 * k + i] = xa[at.amb_off + i];
        for (int i = 0; i < 4; ++i) poses[7 * k + 3 + i] = xa[aq.amb_off + i];
      }
      if (glio_eval_unary(ctx, W, poses.data(), 0, want_jac ? Hd.data() : nullptr, want_jac ? gd.data() : nullptr, cd.data()) != GLIO_OK) return false;
      for (int k = 0; k < W; ++k) {
        ct += cd[k];
        if (!want_jac) continue;
        const int off[2] = {P.active[P.index[slots[k].t]].tan_off, P.active[P.index[slots[k].q]].tan_off};
        for (int p = 0; p < 6; ++p) {
          const int ip = off[p / 3] + p % 3;
          g[ip] += gd[6 * k + p];
          for (int q = 0; q < 6; ++q) { const int iq = off[q / 3] + q % 3; if (ip >= iq) H->at(ip, iq) += Hd[36 * k + 6 * p + q]; }
        }
      }
    }
    if (PB > 0) {
      for (int k = 0; k < KB; ++k) {
        const ActiveBlock& at = P.active[P.index[bkfs[k].first]]; const ActiveBlock& aq = P.active[P.index[bkfs[k].second]];
        for (int i = 0; i < 3; ++i) bposes[7 * k + i] = xa[at.amb_off + i];
        for (int i = 0; i < 4; ++i) bposes[7 * k + 3 + i] = xa[aq.amb_off + i];
      }
      double bc = 0;
      if (glio_eval_binary(ctx, KB, bposes.data(), want_jac ? bHd.data() : nullptr, want_jac ? bHo.data() : nullptr, want_jac ? bg.data() : nullptr, &bc) != GLIO_OK) return false;
      ct += bc;
      if (want_jac) {
        auto toff = [&](int kf, int p) { return (p < 3 ? P.active[P.index[bkfs[kf].first]].tan_off : P.active[P.index[bkfs[kf].second]].tan_off) + p % 3; };
        for (int k = 0; k < KB; ++k)
          for (int p = 0; p < 6; ++p) {
            const int ip = toff(k, p);
            g[ip] += bg[6 * k + p];
            for (int q = 0; q < 6; ++q) { const int iq = toff(k, q); if (ip >= iq) H->at(ip, iq) += bHd[36 * k + 6 * p + q]; }
          }
        for (int pr = 0; pr < PB; ++pr)                       // block (cur, oth): row = cur tangent, col = oth tangent; lower triangle holds it once
          for (int p = 0; p < 6; ++p) for (int q = 0; q < 6; ++q) {
            const int ip = toff(bpairs[pr].kc, p), iq = toff(bpairs[pr].ko, q);
            const double v = bHo[36 * pr + 6 * p + q];
            if (ip >= iq) H->at(ip, iq) += v; else H->at(iq, ip) += v;
          }
      }
    }
    // host blocks (ResidualBlock::Evaluate, ceres.tgz::internal/ceres/residual_block.cc:70-197)
    std::vector<const double*> pp; std::vector<double*> jj; std::vector<std::vector<double>> Jamb, Jloc; std::vector<double> r;
    double fc = 0;
    for (const Problem::Residual* rb : host_blocks) {
      const int nb = (int)rb->params.size(), nr = rb->cost->num_residuals();
      const std::vector<int32_t>& sz = rb->cost->parameter_block_sizes();
      pp.assign(nb, nullptr); jj.assign(nb, nullptr); Jamb.resize(nb); Jloc.resize(nb); r.assign(nr, 0.0);
      bool any_active = false;
      for (int b = 0; b < nb; ++b) {
        auto it = P.index.find(rb->params[b]);
        if (it != P.index.end()) { pp[b] = xa + P.active[it->second].amb_off; any_active = true; if (want_jac) { Jamb[b].assign((size_t)nr * sz[b], 0.0); jj[b] = Jamb[b].data(); } }
        else pp[b] = rb->params[b];
      }
      if (!any_active && fixed_done) continue;           // all-constant block: part of the fixed cost, evaluated once
      if (!rb->cost->Evaluate(pp.data(), r.data(), want_jac && any_active ? jj.data() : nullptr)) return false;
      double sq = 0; for (int k = 0; k < nr; ++k) sq += r[k] * r[k];
      double rho[3] = {sq, 1.0, 0.0};
      if (rb->loss) rb->loss->Evaluate(sq, rho);
      if (!any_active) { fc += 0.5 * rho[0]; continue; }
      ct += 0.5 * rho[0];
      if (!want_jac) continue;
      for (int b = 0; b < nb; ++b) {
        if (!jj[b]) continue;
        const ActiveBlock& a = P.active[P.index[rb->params[b]]];
        if (a.param) {
          std::vector<double> Pm((size_t)a.size * a.local);
          a.param->ComputeJacobian(pp[b], Pm.data());
          Jloc[b].assign((size_t)nr * a.local, 0.0);
          for (int k = 0; k < nr; ++k) for (int c = 0; c < a.local; ++c) { double s = 0; for (int m = 0; m < a.size; ++m) s += Jamb[b][(size_t)k * a.size + m] * Pm[(size_t)m * a.local + c]; Jloc[b][(size_t)k * a.local + c] = s; }
        } else Jloc[b] = Jamb[b];
      }
      if (rb->loss) {
        Corr corr(sq, rho);
        for (int b = 0; b < nb; ++b) if (jj[b]) corr.jac(nr, P.active[P.index[rb->params[b]]].local, r.data(), Jloc[b].data());
        for (int k = 0; k < nr; ++k) r[k] *= corr.residual_scaling;
      }
      for (int a_ = 0; a_ < nb; ++a_) {
        if (!jj[a_]) continue;
        const ActiveBlock& A = P.active[P.index[rb->params[a_]]];
        for (int i = 0; i < A.local; ++i) {
          double gi = 0; for (int k = 0; k < nr; ++k) gi += Jloc[a_][(size_t)k * A.local + i] * r[k];
          g[A.tan_off + i] += gi;
        }
        for (int b_ = 0; b_ < nb; ++b_) {
          if (!jj[b_]) continue;
          const ActiveBlock& B = P.active[P.index[rb->params[b_]]];
          for (int i = 0; i < A.local; ++i) for (int j = 0; j < B.local; ++j) {
            const int ri = A.tan_off + i, cj = B.tan_off + j;
            if (ri < cj) continue;
            double h = 0; for (int k = 0; k < nr; ++k) h += Jloc[a_][(size_t)k * A.local + i] * Jloc[b_][(size_t)k * B.local + j];
            H->at(ri, cj) += h;
          }
        }
      }
    }
    if (!fixed_done) { fixed_cost = fc; fixed_done = true; }
    *cost = ct;
    return std::isfinite(ct);
  };

  glio::SolverSummary gs;
  solver.solve(x.data(), eval, &gs);
  for (const ActiveBlock& a : P.active) std::memcpy(a.user, &x[a.amb_off], sizeof(double) * a.size);
  S.termination_type = gs.termination == glio::TERM_CONVERGENCE ? CONVERGENCE : gs.termination == glio::TERM_NO_CONVERGENCE ? NO_CONVERGENCE : FAILURE;
  S.message += gs.message;
  S.fixed_cost = fixed_cost;
  S.initial_cost = gs.initial_cost + fixed_cost; S.final_cost = gs.final_cost + fixed_cost;
  S.num_successful_steps = gs.num_successful_steps; S.num_unsuccessful_steps = gs.num_unsuccessful_steps;
  for (const glio::IterationRecord& r : gs.iterations) {
    IterationSummary it; it.iteration = r.iteration; it.step_is_valid = r.step_is_valid != 0; it.step_is_successful = r.step_is_successful != 0;
    it.cost = r.cost + fixed_cost; it.cost_change = r.cost_change; it.gradient_max_norm = r.gradient_max_norm; it.gradient_norm = r.gradient_norm;
    it.step_norm = r.step_norm; it.relative_decrease = r.relative_decrease; it.trust_region_radius = r.trust_region_radius;
    S.iterations.push_back(it);
  }
  S.total_time_in_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  if (options.minimizer_progress_to_stdout)
    for (const IterationSummary& it : S.iterations) printf("iter %3d cost %.6e |grad| %.3e radius %.3e %s\n", it.iteration, it.cost, it.gradient_max_norm, it.trust_region_radius, it.step_is_successful ? "" : "(rejected)");
}

inline bool Problem::Evaluate(const EvaluateOptions& eo, double* cost, std::vector<double>* residuals, std::vector<double>* gradient, void*) {
  // host evaluation of every residual block at the current user state (no device routing: this is the slow, exact path
  // the reference never calls; kept for API completeness)
  double ct = 0;
  if (residuals) residuals->clear();
  std::map<double*, int> toff; int nt = 0;
  // every parameter block of the problem owns its columns, constant ones with a zero gradient (Ceres: problem_test.cc:1363-1394)
  for (double* u : order_) { const Block& b = blocks_.at(u); toff[u] = nt; nt += b.param ? b.param->LocalSize() : b.size; }
  if (gradient) gradient->assign(nt, 0.0);
  for (auto& rb : residual_blocks_) {
    const int nb = (int)rb->params.size(), nr = rb->cost->num_residuals();
    const std::vector<int32_t>& sz = rb->cost->parameter_block_sizes();
    std::vector<const double*> pp(rb->params.begin(), rb->params.end());
    std::vector<std::vector<double>> J(nb); std::vector<double*> jj(nb, nullptr); std::vector<double> r(nr);
    if (gradient) for (int b = 0; b < nb; ++b) if (toff.count(rb->params[b]) && !blocks_.at(rb->params[b]).constant) { J[b].assign((size_t)nr * sz[b], 0.0); jj[b] = J[b].data(); }
    if (!rb->cost->Evaluate(pp.data(), r.data(), gradient ? jj.data() : nullptr)) return false;
    double sq = 0; for (double v : r) sq += v * v;
    double rho[3] = {sq, 1.0, 0.0};
    if (rb->loss && eo.apply_loss_function) rb->loss->Evaluate(sq, rho);
    ct += 0.5 * rho[0];
    if (residuals) { shim_internal::Corr c(sq, rho); for (double v : r) residuals->push_back(v * c.residual_scaling); }
    if (gradient) for (int b = 0; b < nb; ++b) {
      if (!jj[b]) continue;
      const Block& blk = blocks_.at(rb->params[b]);
      const int loc = blk.param ? blk.param->LocalSize() : blk.size;
      std::vector<double> Pm;
      if (blk.param) { Pm.resize((size_t)blk.size * loc); blk.param->ComputeJacobian(rb->params[b], Pm.data()); }
      for (int c = 0; c < loc; ++c) {
        double s = 0;
        for (int k = 0; k < nr; ++k) {
          double jl = 0;
          if (blk.param) for (int m = 0; m < blk.size; ++m) jl += J[b][(size_t)k * blk.size + m] * Pm[(size_t)m * loc + c]; else jl = J[b][(size_t)k * blk.size + c];
          s += jl * r[k] * rho[1];
        }
        (*gradient)[toff[rb->params[b]] + c] += s;
      }
    }
  }
  if (cost) *cost = ct;
  return true;
}

}  // namespace ceres
#endif  // GLIO_SHIM_CERES_SHIM_SOLVE_H_
