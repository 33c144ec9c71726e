// Literal drop-in for the reference's CPU build: the reference's own call
//     Theta_map4<W_range>(geom, P, E, B, dt);        (include/strugepic_propagators.hpp:574-583)
// with the AMReX containers left on the host, executed by libstrugepic_b200 on the GPU.
//
// A maintainer of MoPHA/strugepic adds this header to the tree (it includes the reference's own
// strugepic_defs.hpp for CParticleContainer / CParIter and the M, Q, VX.. indices, include/strugepic_defs.hpp:22-49)
// and calls strugepic_b200_amrex::Theta_map4<W>(...) -- or any of the other entry points below -- where the reference's
// functions were called.  Per call: valid cells of E and B -> [comp][k][j][i] host arrays -> spic_set_field; AoS
// particles -> one SoA species per distinct (q, m) -> spic_set_particles; the C-ABI call; and the way back
// (spic_get_field -> Array4, spic_get_particles -> AoS + P.Redistribute()).  That is the path bench.py times as `e2e`.
// Only AMReX calls the reference itself makes are used (MFIter / validbox / array, ParIter / GetArrayOfStructs,
// Geometry::Domain / isPeriodic, FillBoundary, Redistribute; src/strugepic_util.cpp:28-30, 353-354), so the header
// compiles against AMReX and against the stand-in the parity oracle is built with (oracle/amrex_shim).
//
// One rank, ProbLo = 0, dx = 1 (what every shipped driver uses, test/single_particle/main.cpp:106-107).  State that
// lives on the GPU between calls (the bins, the deferred half kick) is rebuilt / flushed per call: this adapter is the
// compatibility path, include/strugepic_b200.hpp (state resident in HBM) is the fast one.
#pragma once
#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "strugepic_b200.h"
#include "strugepic_defs.hpp"  // the reference's: CParticleContainer, CParIter, X Y Z, M Q VX VY VZ

namespace strugepic_b200_amrex {

namespace detail {

inline void check(spic_ctx* ctx, int rc, const char* what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + spic_last_error(ctx));
}

// one context per (box, periodicity, interpolation range), created on first use
inline spic_ctx* context_for(const amrex::Geometry& geom, int w_range) {
  static std::map<std::tuple<int, int, int, int, int, int, int>, spic_ctx*> cache;
  const auto dom = geom.Domain();
  const int n[3] = {dom.length(0), dom.length(1), dom.length(2)};
  const auto key = std::make_tuple(n[0], n[1], n[2], (int)geom.isPeriodic(0), (int)geom.isPeriodic(1),
                                   (int)geom.isPeriodic(2), w_range);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  spic_config cfg{};
  for (int d = 0; d < 3; ++d) {
    cfg.n_cell[d] = n[d];
    cfg.periodic[d] = geom.isPeriodic(d) ? 1 : 0;
  }
  cfg.interp = w_range == 2 ? SPIC_INTERP_P8R2 : SPIC_INTERP_PWL;
  cfg.engine = SPIC_ENGINE_BINNED;
  cfg.nranks = 1;
  spic_ctx* ctx = nullptr;
  const int rc = spic_create(&cfg, &ctx);
  if (rc) throw std::runtime_error(std::string("spic_create: ") + spic_last_error(nullptr));
  cache[key] = ctx;
  return ctx;
}

inline void upload_field(spic_ctx* ctx, const amrex::Geometry& geom, amrex::MultiFab& F, int which,
                         std::vector<double>& h) {
  const auto dom = geom.Domain();
  const long nx = dom.length(0), ny = dom.length(1), nz = dom.length(2);
  h.resize((size_t)(3 * nx * ny * nz));
  for (amrex::MFIter mfi(F); mfi.isValid(); ++mfi) {
    const amrex::Box& box = mfi.validbox();
    amrex::Array4<amrex::Real> const& a = F.array(mfi);
    const auto lo = box.loVect();
    const auto hi = box.hiVect();
    for (int c = 0; c < 3; ++c)
      for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i) h[(size_t)(((c * nz + k) * ny + j) * nx + i)] = a(i, j, k, c);
  }
  check(ctx, spic_set_field(ctx, which, h.data()), "spic_set_field");
}

inline void download_field(spic_ctx* ctx, const amrex::Geometry& geom, amrex::MultiFab& F, int which,
                           std::vector<double>& h) {
  const auto dom = geom.Domain();
  const long nx = dom.length(0), ny = dom.length(1), nz = dom.length(2);
  h.resize((size_t)(3 * nx * ny * nz));
  check(ctx, spic_get_field(ctx, which, h.data()), "spic_get_field");
  for (amrex::MFIter mfi(F); mfi.isValid(); ++mfi) {
    const amrex::Box& box = mfi.validbox();
    amrex::Array4<amrex::Real> const& a = F.array(mfi);
    const auto lo = box.loVect();
    const auto hi = box.hiVect();
    for (int c = 0; c < 3; ++c)
      for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i) a(i, j, k, c) = h[(size_t)(((c * nz + k) * ny + j) * nx + i)];
  }
  F.FillBoundary(geom.periodicity());  // the guards the reference's next consumer would find filled
}

struct SpeciesBuf {
  double q, m;
  std::vector<double> a[6];
};

// AoS particles -> one SoA species per distinct (q, m) (the reference carries q and m per particle; the kernels hoist
// them per species).  Species are created on the first call and refilled afterwards.
inline void upload_particles(spic_ctx* ctx, CParticleContainer& P, std::vector<SpeciesBuf>& sp) {
  for (auto& s : sp)
    for (auto& t : s.a) t.clear();
  std::map<std::pair<double, double>, size_t> index;
  for (size_t i = 0; i < sp.size(); ++i) index[std::make_pair(sp[i].q, sp[i].m)] = i;
  for (CParIter pti(P, 0); pti.isValid(); ++pti) {
    auto& particles = pti.GetArrayOfStructs();
    const long np = pti.numParticles();
    for (long n = 0; n < np; ++n) {
      auto& p = particles[n];
      const auto key = std::make_pair((double)p.rdata(Q), (double)p.rdata(M));
      auto it = index.find(key);
      if (it == index.end()) {
        sp.push_back(SpeciesBuf{key.first, key.second, {}});
        it = index.emplace(key, sp.size() - 1).first;
      }
      SpeciesBuf& s = sp[it->second];
      s.a[0].push_back(p.pos(X));
      s.a[1].push_back(p.pos(Y));
      s.a[2].push_back(p.pos(Z));
      s.a[3].push_back(p.rdata(VX));
      s.a[4].push_back(p.rdata(VY));
      s.a[5].push_back(p.rdata(VZ));
    }
  }
  const int have = spic_num_species(ctx);
  for (size_t i = 0; i < sp.size(); ++i) {
    SpeciesBuf& s = sp[i];
    const int64_t n = (int64_t)s.a[0].size();
    if ((int)i < have)
      check(ctx, spic_set_particles(ctx, (int)i, n, s.a[0].data(), s.a[1].data(), s.a[2].data(), s.a[3].data(),
                                    s.a[4].data(), s.a[5].data()), "spic_set_particles");
    else
      check(ctx, spic_add_species(ctx, s.q, s.m, n, s.a[0].data(), s.a[1].data(), s.a[2].data(), s.a[3].data(),
                                  s.a[4].data(), s.a[5].data()), "spic_add_species");
  }
}

// SoA species -> the container's AoS (the particle order is the library's cell order: the reference's own
// Redistribute reorders too, and nothing in it depends on the order)
inline void download_particles(spic_ctx* ctx, CParticleContainer& P, std::vector<SpeciesBuf>& sp) {
  size_t total = 0;
  for (size_t i = 0; i < sp.size(); ++i) {
    int64_t n = 0;
    check(ctx, spic_num_particles(ctx, (int)i, &n), "spic_num_particles");
    for (auto& t : sp[i].a) t.resize((size_t)n);
    check(ctx, spic_get_particles(ctx, (int)i, sp[i].a[0].data(), sp[i].a[1].data(), sp[i].a[2].data(),
                                  sp[i].a[3].data(), sp[i].a[4].data(), sp[i].a[5].data()), "spic_get_particles");
    total += (size_t)n;
  }
  // refill the tiles in place: every particle slot of the container takes the next particle of the species lists
  size_t si = 0, pi = 0;
  auto next = [&](CParticle& p) {
    while (si < sp.size() && pi >= sp[si].a[0].size()) {
      ++si;
      pi = 0;
    }
    if (si >= sp.size()) return false;
    const SpeciesBuf& s = sp[si];
    p.pos(X) = s.a[0][pi];
    p.pos(Y) = s.a[1][pi];
    p.pos(Z) = s.a[2][pi];
    p.rdata(M) = s.m;
    p.rdata(Q) = s.q;
    p.rdata(VX) = s.a[3][pi];
    p.rdata(VY) = s.a[4][pi];
    p.rdata(VZ) = s.a[5][pi];
    ++pi;
    return true;
  };
  size_t filled = 0;
  for (CParIter pti(P, 0); pti.isValid(); ++pti) {
    auto& particles = pti.GetArrayOfStructs();
    const long np = pti.numParticles();
    for (long n = 0; n < np; ++n)
      if (next(particles[n])) ++filled;
  }
  if (filled != total) throw std::runtime_error("strugepic_b200_amrex: the particle count changed inside a map");
  P.Redistribute();  // positions are valid; every particle finds its grid / tile again (hpp:368)
}

struct State {
  std::vector<double> hostE, hostB;
  std::vector<SpeciesBuf> species;
};
inline State& state_of(spic_ctx* ctx) {
  static std::map<spic_ctx*, State> m;
  return m[ctx];
}

template <class Op>
inline void run(const amrex::Geometry& geom, CParticleContainer& P, amrex::MultiFab& E, amrex::MultiFab& B, int w_range,
                Op op) {
  spic_ctx* ctx = context_for(geom, w_range);
  State& st = state_of(ctx);
  upload_field(ctx, geom, E, SPIC_FIELD_E, st.hostE);
  upload_field(ctx, geom, B, SPIC_FIELD_B, st.hostB);
  upload_particles(ctx, P, st.species);
  check(ctx, op(ctx), "strugepic_b200 sub-flow");
  download_field(ctx, geom, E, SPIC_FIELD_E, st.hostE);
  download_field(ctx, geom, B, SPIC_FIELD_B, st.hostB);
  download_particles(ctx, P, st.species);
}

}  // namespace detail

// ---- the reference's entry points, same signatures (hpp:548-583, 347-372, 52-71; cpp:102-113) -------------------
template <int W_range>
inline void Theta_map1(const amrex::Geometry geom, CParticleContainer& P, amrex::MultiFab& E, amrex::MultiFab& B, double dt) {
  detail::run(geom, P, E, B, W_range, [&](spic_ctx* c) { return spic_map(c, 1, dt); });
}
template <int W_range>
inline void Theta_map2(const amrex::Geometry geom, CParticleContainer& P, amrex::MultiFab& E, amrex::MultiFab& B, double dt) {
  detail::run(geom, P, E, B, W_range, [&](spic_ctx* c) { return spic_map(c, 2, dt); });
}
template <int W_range>
inline void Theta_map4(const amrex::Geometry geom, CParticleContainer& P, amrex::MultiFab& E, amrex::MultiFab& B, double dt) {
  detail::run(geom, P, E, B, W_range, [&](spic_ctx* c) { return spic_map(c, 4, dt); });
}
template <int comp, int W_range>
inline void G_Theta(const amrex::Geometry geom, CParticleContainer& P, amrex::MultiFab& E, amrex::MultiFab& B, double dt) {
  detail::run(geom, P, E, B, W_range, [&](spic_ctx* c) { return spic_theta_axis(c, comp, dt); });
}
template <int W_range>
inline void G_Theta_E(const amrex::Geometry geom, CParticleContainer& P, amrex::MultiFab& E, amrex::MultiFab& B, double dt) {
  detail::run(geom, P, E, B, W_range, [&](spic_ctx* c) { return spic_theta_E(c, dt); });
}
template <int W_range>
inline void G_Theta_B(const amrex::Geometry geom, CParticleContainer& P, amrex::MultiFab& E, amrex::MultiFab& B, double dt) {
  detail::run(geom, P, E, B, W_range, [&](spic_ctx* c) { return spic_theta_B(c, dt); });
}

}  // namespace strugepic_b200_amrex
