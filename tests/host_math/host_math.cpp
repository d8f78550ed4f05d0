// Host-side check of the kernel arithmetic (TEST ONLY — never part of the product library).
// Compiles adsurf_b200/csrc/adsurf_math.cuh, the exact per-sample code the CUDA kernels run,
// for the CPU so tests/test_host_math.py can verify the hand-derived adjoint against the oracle
// without a GPU.  Built on the fly by the test with g++.
#include <vector>

#include "adsurf_math.cuh"
#include "adsurf_dispersion.cuh"

using namespace adsurf;

namespace {
struct HostSink {
  double *gd, *ga, *gb, *gr;  // accumulate [L]
  double *jd, *ja, *jb, *jr;  // per-sample rows [L] or null
  void operator()(int m, double d, double a, double b, double r) const {
    gd[m] += d; ga[m] += a; gb[m] += b; gr[m] += r;
    if (jd) { jd[m] = d; ja[m] = a; jb[m] = b; jr[m] = r; }
  }
};
}  // namespace

extern "C" {

// modal dispersion curves of S models by the kernel's root search (adsurf_dispersion.cuh), run on the host
void host_dispersion(const double* d, const double* a, const double* b, const double* rho, const double* t, int S, int L,
                     int NF, int nmodes, double dc, int ifunc, double* c, int* status) {
  std::vector<double> work(NF);
  for (int s = 0; s < S; ++s) {
    const DispModel M{d + (size_t)s * L, a + (size_t)s * L, b + (size_t)s * L, rho + (size_t)s * L, L, ifunc};
    DispScalarEval ev{M, 0.0};
    status[s] = disp_getc(ev, M, t + (size_t)s * NF, NF, nmodes, dc, c + (size_t)s * nmodes * NF, work.data());
  }
}

// elementary functions of the kernel header, for accuracy tests against libm
void host_elementary(const double* x, int n, double* rs, double* ex, double* sn, double* cs) {
  for (int i = 0; i < n; ++i) {
    rs[i] = rsqrt_hd(x[i]);
    ex[i] = exp_neg(x[i]);
    sincos_fast(x[i], sn + i, cs + i);
  }
}

// F[S,N], gradients reduced per model with seeds gF (null = 1), gc[S,N], optional Jacobians [S,N,L]
int host_points(const double* d, const double* alpha, const double* beta, const double* rho, const double* t,
                const double* c, const double* gF, int S, int L, int N, double* F, double* gd, double* ga, double* gb,
                double* grho, double* gc, double* Jd, double* Ja, double* Jb, double* Jrho) {
  std::vector<double> par(par_doubles(L)), est(kSlot * (L > 1 ? L - 1 : 1));
  for (int s = 0; s < S; ++s) {
    for (int m = 0; m < L; ++m)
      stage_one(par.data(), L, m, d[s * L + m], alpha[s * L + m], beta[s * L + m], rho[s * L + m]);
    for (int m = 0; m < L; ++m) gd[s * L + m] = ga[s * L + m] = gb[s * L + m] = grho[s * L + m] = 0.0;
    for (int i = 0; i < N; ++i) {
      const size_t idx = (size_t)s * N + i;
      const double ic = 1.0 / c[idx];
      const Sample z = make_sample(omega_of_period(t[idx]), ic);
      const double v = sample_forward_store(DirectSource{par.data(), L}, z, PlainStore<1>{est.data()});
      const double v2 = sample_forward(DirectSource{par.data(), L}, z);
      if (v != v2 && !(v != v && v2 != v2)) return -1;
      F[idx] = v;
      HostSink sink{gd + s * L, ga + s * L, gb + s * L, grho + s * L,
                    Jd ? Jd + idx * L : nullptr, Ja ? Ja + idx * L : nullptr, Jb ? Jb + idx * L : nullptr,
                    Jrho ? Jrho + idx * L : nullptr};
      const double gic = sample_adjoint(DirectSource{par.data(), L}, z, gF ? gF[idx] : 1.0, v,
                                        PlainStore<1>{est.data()}, sink);
      if (gc) gc[idx] = -ic * z.ic * gic;  // d(k/omega)/dc
    }
  }
  return 0;
}

// det[S,K,NF] and gradients reduced per model with seeds gdet (null = 1)
int host_grid(const double* d, const double* alpha, const double* beta, const double* rho, const double* t,
              const double* c, const double* gdet, int S, int L, int NF, int K, double* det, double* gd, double* ga,
              double* gb, double* grho) {
  std::vector<double> par(par_doubles(L)), est(kSlot * (L > 1 ? L - 1 : 1));
  for (int s = 0; s < S; ++s) {
    for (int m = 0; m < L; ++m)
      stage_one(par.data(), L, m, d[s * L + m], alpha[s * L + m], beta[s * L + m], rho[s * L + m]);
    stage_bounds(par.data(), L, d + s * L, beta + s * L);
    for (int m = 0; m < L; ++m) gd[s * L + m] = ga[s * L + m] = gb[s * L + m] = grho[s * L + m] = 0.0;
    // dense path: per-(velocity, layer) table (sentinel, layers, sentinel) + per-period lanes, as the dense kernels do
    std::vector<double> tab((size_t)table_doubles(L, 1, kTabAdj) + 2), tabf((size_t)table_doubles(L, 1, kTabFwd) + 2);
    double* tb = tab.data() + (((size_t)tab.data() & 15) ? 1 : 0);  // 16-byte aligned like the shared table
    double* tf = tabf.data() + (((size_t)tabf.data() & 15) ? 1 : 0);
    std::vector<double> estd(kSlot * (size_t)L);  // one spare slot: the dense read-ahead touches it
    for (int k = 0; k < K; ++k) {
      const double ic = 1.0 / c[s * K + k];
      table_sentinel(tb, kTabAdj);
      table_sentinel(tf, kTabFwd);  // the forward-only kernel reads 16-double entries
      for (int m = 0; m < L - 1; ++m) {
        table_entry<kTabAdj>(tb + (size_t)(m + 1) * kTabAdj, par.data(), m, ic);
        table_entry<kTabFwd>(tf + (size_t)(m + 1) * kTabFwd, par.data(), m, ic);
      }
      table_sentinel(tb + (size_t)L * kTabAdj, kTabAdj);
      table_sentinel(tf + (size_t)L * kTabFwd, kTabFwd);
      const TableSource<kTabAdj> src{par.data(), L, tb + kTabAdj, nullptr};
      const TableSource<kTabFwd> srcf{par.data(), L, tf + kTabFwd, nullptr};
      for (int f = 0; f < NF; ++f) {
        const size_t idx = ((size_t)s * K + k) * NF + f;
        const double om0 = omega_of_period(t[s * NF + f]);
        const Sample z = make_sample_dense(om0, ic, 1.0 / fmax(om0, 1.0e-4));
        det[idx] = sample_forward_store(src, z, PlainStore<1>{estd.data()});
        const double v2 = sample_forward(srcf, z);
        if (det[idx] != v2 && !(det[idx] != det[idx] && v2 != v2)) return -1;
        HostSink sink{gd + s * L, ga + s * L, gb + s * L, grho + s * L, nullptr, nullptr, nullptr, nullptr};
        WalkingSink<HostSink> walk(sink);
        sample_adjoint(src, z, gdet ? gdet[idx] : 1.0, det[idx], PlainStore<1>{estd.data()}, walk);
      }
    }
  }
  return 0;
}
}
