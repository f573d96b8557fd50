// ref_opengr.cpp -- ORACLE/_ref (TEST INFRASTRUCTURE ONLY).
//
// Drives the reference's OWN Super4PCS matcher (the header-only OpenGR fork under
// /root/reference/src/OpenGR_4pcs/src/gr, compiled where it lies by oracle/build_ref.sh) the way
// pcl::Super4PCS::computeTransformation does (demos/PCLWrapper/pcl/registration/impl/super4pcs.hpp:67-115):
//     MatcherType = gr::Match4pcsBase<gr::FunctorSuper4PCS, Visitor, gr::AdaptivePointFilter, gr::AdaptivePointFilter::Options>
//     matcher._ppfs = ppfs;  matcher.ComputeTransformation(P = scene, Q = model, ...);  -> _pose_hypo, _pose_lcp_scores
// with the options PoseEstimator::runSuper4pcs sets (src/perception/src/PoseEstimator.cpp:62-79).
//
// The subclass below re-states only the ~20-line trial loop (ComputeTransformation / Perform_N_steps / TryOneBase,
// congruentSetExplorationBase.hpp:71-125,129-194,201-216) so that every trial's base and congruent set can be
// recorded next to the hypotheses the reference's TryCongruentSet (:221-340) + Verify (:346-435) +
// ComputeRigidTransformation (matchBase.hpp:230-377) emit for them; those three run unmodified.
// Run single-threaded (OMP_NUM_THREADS=1 or nthreads=1) the emission order is (trial, congruent-set index).
#include <omp.h>

#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include <gr/shared.h>
#include <gr/sampling.h>
#include <gr/utils/logger.h>
#include <gr/algorithms/PointPairFilter.h>
#include <gr/algorithms/match4pcsBase.h>
#include <gr/algorithms/FunctorSuper4pcs.h>

namespace {

struct Visitor {
  template <typename Derived>
  inline void operator()(float, float, const Eigen::MatrixBase<Derived> &) const {}
  constexpr bool needsGlobalTransformation() const { return false; }
};

using MatcherBase = gr::Match4pcsBase<gr::FunctorSuper4PCS, Visitor, gr::AdaptivePointFilter, gr::AdaptivePointFilter::Options>;

struct TrialRecord {
  std::array<int, 4> base;
  int ok;              // generateCongruents succeeded
  int quad_begin, quad_end;
  int hyp_begin, hyp_end;
  // from the restated generateCongruents (match4pcsBase.hpp:207-281): what SelectQuadrilateral and the two ExtractPairs returned
  int base_ok = 0;     // SelectQuadrilateral succeeded
  float inv1 = 0, inv2 = 0, dist1 = 0, dist2 = 0;
  int pairs1_begin = 0, pairs1_end = 0, pairs2_begin = 0, pairs2_end = 0;
};

class Matcher : public MatcherBase {
 public:
  using MatcherBase::MatcherBase;
  std::vector<TrialRecord> trials;
  std::vector<std::array<int, 4>> quads;
  // per quadrilateral, from the reference's own ComputeRigidTransformation / Verify called the way TryCongruentSet does
  std::vector<float> quad_rms, quad_lcp;
  std::vector<int> quad_ok;
  std::vector<std::pair<int, int>> all_pairs;  // pairs1 then pairs2 of every trial (see TrialRecord)

  // generateCongruents (match4pcsBase.hpp:207-281) restated call for call so the intermediate results can be recorded;
  // SelectQuadrilateral, ExtractPairs and FindCongruentQuadrilaterals themselves run unmodified.
  bool generateCongruentsRecorded(CongruentBaseType &base, Set &congruent_quads, TrialRecord &tr) {
    Scalar invariant1, invariant2;
    if (!this->SelectQuadrilateral(invariant1, invariant2, base[0], base[1], base[2], base[3])) return false;
    tr.base_ok = 1; tr.base = base; tr.inv1 = invariant1; tr.inv2 = invariant2;
    const auto &b0 = this->base_3D_[0]; const auto &b1 = this->base_3D_[1];
    const auto &b2 = this->base_3D_[2]; const auto &b3 = this->base_3D_[3];
    const Scalar distance1 = (b0.pos() - b1.pos()).norm();
    const Scalar distance2 = (b2.pos() - b3.pos()).norm();
    tr.dist1 = distance1; tr.dist2 = distance2;
    std::vector<std::pair<int, int>> pairs1, pairs2;
    const Scalar normal_angle1 = (b0.normal() - b1.normal()).norm();
    const Scalar normal_angle2 = (b2.normal() - b3.normal()).norm();
    this->fun_.ExtractPairs(distance1, normal_angle1, this->distance_factor * this->options_.delta, 0, 1, pairs1);
    this->fun_.ExtractPairs(distance2, normal_angle2, this->distance_factor * this->options_.delta, 2, 3, pairs2);
    tr.pairs1_begin = (int)all_pairs.size(); all_pairs.insert(all_pairs.end(), pairs1.begin(), pairs1.end()); tr.pairs1_end = (int)all_pairs.size();
    tr.pairs2_begin = (int)all_pairs.size(); all_pairs.insert(all_pairs.end(), pairs2.begin(), pairs2.end()); tr.pairs2_end = (int)all_pairs.size();
    if (pairs1.size() == 0 || pairs2.size() == 0) return false;
    return this->fun_.FindCongruentQuadrilaterals(invariant1, invariant2, this->distance_factor * this->options_.delta,
                                                  this->distance_factor * this->options_.delta, pairs1, pairs2, &congruent_quads);
  }

  void evalQuads(const CongruentBaseType &base, const Set &set) {
    Coordinates references;
    for (int i = 0; i < 4; ++i) references[i] = this->sampled_P_3D_[base[i]];
    Eigen::Matrix<Scalar, 3, 1> centroid1 = (references[0].pos() + references[1].pos() + references[2].pos()) / Scalar(3);
    for (size_t i = 0; i < set.size(); ++i) {
      Coordinates cand;
      for (int j = 0; j < 4; ++j) cand[j] = this->sampled_Q_3D_[set[i][j]];
      Eigen::Matrix<Scalar, 4, 4> transform;
      Eigen::Matrix<Scalar, 3, 1> centroid2 = (cand[0].pos() + cand[1].pos() + cand[2].pos()) / Scalar(3.);
      Scalar rms = -1;
      const bool ok = this->ComputeRigidTransformation(references, cand, centroid1, centroid2, transform, rms, false);
      Scalar lcp = 0;
      if (ok && rms >= Scalar(0.) && rms < this->options_.delta) lcp = this->Verify(transform);
      quad_ok.push_back(ok ? 1 : 0); quad_rms.push_back(rms); quad_lcp.push_back(lcp);
    }
  }

  void run(const std::vector<gr::Point3D> &P, const std::vector<gr::Point3D> &Q, int max_trials) {
    Visitor v;
    gr::UniformDistSampler sampler;
    // ComputeTransformation (:71-125): number_of_trials_ is evaluated BEFORE init() from the still-uninitialised
    // diameters and always clamps to kMinNumberOfTrials = 30 (SURVEY.md quick facts); restated as the constant.
    this->number_of_trials_ = max_trials > 0 ? max_trials : 30;
    this->current_trial_ = 0;
    this->best_LCP_ = 0.0;
    for (int i = 0; i < 4; ++i) { this->base_[i] = 0; this->current_congruent_[i] = 0; }
    this->init(P, Q, sampler);
    int successes = 0;
    // Perform_N_steps (:129-194) without the wall-clock stop (max_time_seconds): deterministic
    for (int i = 0; i < this->number_of_trials_; ++i) {
      TrialRecord tr;
      tr.base = {0, 0, 0, 0};
      tr.quad_begin = tr.quad_end = (int)quads.size();
      tr.hyp_begin = tr.hyp_end = (int)this->_pose_hypo.size();
      CongruentBaseType base;
      Set congruent;
      tr.ok = this->generateCongruentsRecorded(base, congruent, tr) ? 1 : 0;  // TryOneBase (:201-216)
      if (tr.ok) {
        ++successes;
        tr.base = base;
        for (const auto &q : congruent) quads.push_back(q);
        tr.quad_end = (int)quads.size();
        size_t nb = 0;
        this->TryCongruentSet(base, congruent, v, nb);
        evalQuads(base, congruent);
        tr.hyp_end = (int)this->_pose_hypo.size();
      }
      trials.push_back(tr);
      if (i > this->number_of_trials_ || successes >= this->options_.success_quadrilaterals) break;
    }
  }
  const std::vector<gr::Point3D> &sampledP() const { return this->sampled_P_3D_; }
  const std::vector<gr::Point3D> &sampledQ() const { return this->sampled_Q_3D_; }
  Eigen::Vector3f centroidP() const { return this->centroid_P_; }
  Eigen::Vector3f centroidQ() const { return this->centroid_Q_; }
  float diameter() const { return this->P_diameter_; }
};

void fill(const float *xyz, const float *nrm, const float *prob, int n, std::vector<gr::Point3D> &out) {
  out.clear();
  out.reserve(n);
  for (int i = 0; i < n; ++i) {  // fillPointSet (impl/super4pcs.hpp:87-105)
    out.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    out[i].set_rgb(Eigen::Vector3f(0, 0, 0));
    out[i].set_normal(Eigen::Vector3f(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
    out[i].setProb(prob ? prob[i] : 1.f);
  }
}

Matcher *g_last = nullptr;

}  // namespace

extern "C" {

struct hop_ref_s4pcs_options {
  int32_t sample_size;             // super4pcs_sample_size (100)
  float overlap;                   // super4pcs_overlap (0.2)
  float delta;                     // super4pcs_delta (0.003)
  float dispersion;                // super4pcs_dispersion (0.5)
  int32_t success_quadrilaterals;  // super4pcs_success_quadrilaterals (10)
  float max_normal_difference;     // -1
  float max_color_distance;        // -1
  int32_t max_trials;              // 0 = the reference's effective 30
  int32_t nthreads;                // OpenMP threads for TryCongruentSet (1 = canonical order)
};

// Runs the matcher.  Returns the number of hypotheses (or -1).  Results are fetched with hop_ref_s4pcs_get.
int hop_ref_s4pcs_run(const float *P_xyz, const float *P_nrm, const float *P_prob, int nP, const float *Q_xyz, const float *Q_nrm,
                      int nQ, const int32_t *ppf_keys, int n_keys, const hop_ref_s4pcs_options *o) {
  delete g_last;
  g_last = nullptr;
  omp_set_num_threads(o->nthreads > 0 ? o->nthreads : 1);
  MatcherBase::OptionsType opt;
  opt.sample_size = o->sample_size;                    // PoseEstimator.cpp:66-73
  opt.configureOverlap(o->overlap);
  opt.max_time_seconds = 1000000;                      // the wall-clock stop is disabled in the oracle
  opt.delta = o->delta;
  opt.sample_dispersion = o->dispersion;
  opt.success_quadrilaterals = o->success_quadrilaterals;
  opt.max_normal_difference = o->max_normal_difference;
  opt.max_color_distance = o->max_color_distance;
  static gr::Utils::Logger logger(gr::Utils::NoLog);
  Matcher *m = new Matcher(opt, logger);
  for (int k = 0; k < n_keys; ++k) {
    std::vector<int> key(ppf_keys + 4 * k, ppf_keys + 4 * k + 4);
    m->_ppfs[key];  // only key membership is ever queried (matchBase.hpp:134,159,201)
  }
  std::vector<gr::Point3D> P, Q;
  fill(P_xyz, P_nrm, P_prob, nP, P);
  fill(Q_xyz, Q_nrm, nullptr, nQ, Q);
  m->run(P, Q, o->max_trials);
  g_last = m;
  return (int)m->_pose_hypo.size();
}

// sizes: [0] hypotheses [1] trials [2] quads [3] sampled P [4] sampled Q
void hop_ref_s4pcs_sizes(int32_t *sizes) {
  if (!g_last) { std::memset(sizes, 0, 5 * sizeof(int32_t)); return; }
  sizes[0] = (int32_t)g_last->_pose_hypo.size(); sizes[1] = (int32_t)g_last->trials.size(); sizes[2] = (int32_t)g_last->quads.size();
  sizes[3] = (int32_t)g_last->sampledP().size(); sizes[4] = (int32_t)g_last->sampledQ().size();
}

// poses: H x 16 column-major (global frame); lcp: H; trials: T x 9 (base[4], ok, quad_begin, quad_end, hyp_begin, hyp_end);
// quads: M x 4; Pc / Qc: centred sampled clouds (n x 3) + normals; centroids: 6 floats (P then Q); misc[0] = diameter
void hop_ref_s4pcs_get(float *poses, float *lcp, int32_t *trials, int32_t *quads, float *Pc, float *Pn, float *Qc, float *Qn,
                       float *centroids, float *misc) {
  if (!g_last) return;
  const Matcher &m = *g_last;
  for (size_t i = 0; i < m._pose_hypo.size(); ++i) {
    std::memcpy(poses + 16 * i, m._pose_hypo[i].data(), 16 * sizeof(float));
    lcp[i] = m._pose_lcp_scores[i];
  }
  for (size_t i = 0; i < m.trials.size(); ++i) {
    const TrialRecord &t = m.trials[i];
    int32_t *r = trials + 9 * i;
    for (int k = 0; k < 4; ++k) r[k] = t.base[k];
    r[4] = t.ok; r[5] = t.quad_begin; r[6] = t.quad_end; r[7] = t.hyp_begin; r[8] = t.hyp_end;
  }
  for (size_t i = 0; i < m.quads.size(); ++i)
    for (int k = 0; k < 4; ++k) quads[4 * i + k] = m.quads[i][k];
  auto dump = [](const std::vector<gr::Point3D> &c, float *xyz, float *nrm) {
    for (size_t i = 0; i < c.size(); ++i)
      for (int k = 0; k < 3; ++k) { xyz[3 * i + k] = c[i].pos()[k]; nrm[3 * i + k] = c[i].normal()[k]; }
  };
  dump(m.sampledP(), Pc, Pn);
  dump(m.sampledQ(), Qc, Qn);
  Eigen::Vector3f cp = m.centroidP(), cq = m.centroidQ();
  for (int k = 0; k < 3; ++k) { centroids[k] = cp[k]; centroids[3 + k] = cq[k]; }
  misc[0] = m.diameter();
}

// per trial (T x 9 ints): base_ok, base[4], pairs1_begin, pairs1_end, pairs2_begin, pairs2_end; (T x 4 floats): inv1, inv2, dist1, dist2;
// pairs: n_pairs x 2 ints (hop_ref_s4pcs_num_pairs of them)
int hop_ref_s4pcs_num_pairs() { return g_last ? (int)g_last->all_pairs.size() : 0; }
void hop_ref_s4pcs_get_trials(int32_t *ti, float *tf, int32_t *pairs) {
  if (!g_last) return;
  for (size_t i = 0; i < g_last->trials.size(); ++i) {
    const TrialRecord &t = g_last->trials[i];
    int32_t *r = ti + 9 * i;
    r[0] = t.base_ok; for (int k = 0; k < 4; ++k) r[1 + k] = t.base[k];
    r[5] = t.pairs1_begin; r[6] = t.pairs1_end; r[7] = t.pairs2_begin; r[8] = t.pairs2_end;
    tf[4 * i] = t.inv1; tf[4 * i + 1] = t.inv2; tf[4 * i + 2] = t.dist1; tf[4 * i + 3] = t.dist2;
  }
  for (size_t i = 0; i < g_last->all_pairs.size(); ++i) { pairs[2 * i] = g_last->all_pairs[i].first; pairs[2 * i + 1] = g_last->all_pairs[i].second; }
}

// per quadrilateral: ok flag, rms, lcp (0 when gated out)
void hop_ref_s4pcs_get_quads(int32_t *ok, float *rms, float *lcp) {
  if (!g_last) return;
  for (size_t i = 0; i < g_last->quad_ok.size(); ++i) { ok[i] = g_last->quad_ok[i]; rms[i] = g_last->quad_rms[i]; lcp[i] = g_last->quad_lcp[i]; }
}

// gr::computePPF (matchBase.hpp:47-68) for one pair: the integer key the table is probed with
void hop_ref_compute_ppf(const float *p1, const float *n1, const float *p2, const float *n2, int32_t *key) {
  gr::Point3D a(p1[0], p1[1], p1[2]), b(p2[0], p2[1], p2[2]);
  a.set_normal(Eigen::Vector3f(n1[0], n1[1], n1[2]));
  b.set_normal(Eigen::Vector3f(n2[0], n2[1], n2[2]));
  std::vector<int> ppf;
  gr::computePPF(a, b, ppf);
  for (int k = 0; k < 4; ++k) key[k] = ppf[k];
}

// the table computePPF.cpp:86-98 builds: keys of all pairs (i < j) of the cloud; keys_out: n*(n-1)/2 x 4
void hop_ref_ppf_pairs(const float *xyz, const float *nrm, int n, int32_t *keys_out) {
  std::vector<gr::Point3D> pts;
  fill(xyz, nrm, nullptr, n, pts);
  size_t k = 0;
  std::vector<int> ppf;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j, ++k) {
      gr::computePPF(pts[i], pts[j], ppf);
      for (int c = 0; c < 4; ++c) keys_out[4 * k + c] = ppf[c];
    }
}

}  // extern "C"
