// chain_host.hpp — host-side construction of ChainP (canonical link frames) from the flat URDF-style description in
// ilqr_problem.chain (include/ilqr_b200.h).  Used by ilqr_create (capi.cu) and by the host compile of the device
// linearisation in tests/.
#pragma once
#include <cmath>
#include <cstring>

#include "../../include/ilqr_b200.h"
#include "chain_params.cuh"

namespace ilqr {

inline void build_chain_params(const ilqr_problem& p, bool floating, ChainP& c) {
  c.nq = p.nq; c.dt = p.dt;
  for (int k = 0; k < 3; ++k) c.g[k] = p.gravity[k];
  if (floating) {   // root link inertial (row nq), in the base frame, which is left as it is
    const double* r = p.chain + p.nq * ILQR_CHAIN_STRIDE;
    c.base_mass = r[9];
    for (int k = 0; k < 3; ++k) c.base_com[k] = r[10 + k];
    for (int k = 0; k < 6; ++k) c.base_I[k] = r[13 + k];
  }
  // Canonical link frames: L'_i = L_i·C_i with C_i ẑ = axis_i, so every joint turns about its own +z
  // (chain.cuh).  x_{L'_{i-1}} = C_{i-1}ᵀ xyz_i + (C_{i-1}ᵀ R0_i C_i)·Rot(z, q_i)·x_{L'_i}.
  double Cprev[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // row-major; the base frame is left alone
  for (int i = 0; i < p.nq; ++i) {
    const double* r = p.chain + i * ILQR_CHAIN_STRIDE;
    const double ax[3] = {r[6], r[7], r[8]};
    double C[9];
    if (ax[2] == 1.0 && ax[0] == 0.0 && ax[1] == 0.0) { const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; std::memcpy(C, I3, sizeof C); }
    else if (ax[0] == 1.0 && ax[1] == 0.0 && ax[2] == 0.0) { const double P[9] = {0, 0, 1, 1, 0, 0, 0, 1, 0}; std::memcpy(C, P, sizeof C); }  // columns (ŷ, ẑ, x̂)
    else if (ax[1] == 1.0 && ax[0] == 0.0 && ax[2] == 0.0) { const double P[9] = {0, 1, 0, 0, 0, 1, 1, 0, 0}; std::memcpy(C, P, sizeof C); }  // columns (ẑ, x̂, ŷ)
    else {
      // columns (x', y', a): x' ⟂ a from the basis vector least aligned with a, y' = a × x'
      int kmin = 0;
      for (int k = 1; k < 3; ++k) if (std::fabs(ax[k]) < std::fabs(ax[kmin])) kmin = k;
      double hx[3] = {0, 0, 0}; hx[kmin] = 1.0;
      const double d = ax[kmin];
      double xp[3] = {hx[0] - d * ax[0], hx[1] - d * ax[1], hx[2] - d * ax[2]};
      const double nrm = std::sqrt(xp[0] * xp[0] + xp[1] * xp[1] + xp[2] * xp[2]);
      for (int k = 0; k < 3; ++k) xp[k] /= nrm;
      const double yp[3] = {ax[1] * xp[2] - ax[2] * xp[1], ax[2] * xp[0] - ax[0] * xp[2], ax[0] * xp[1] - ax[1] * xp[0]};
      for (int k = 0; k < 3; ++k) { C[3 * k + 0] = xp[k]; C[3 * k + 1] = yp[k]; C[3 * k + 2] = ax[k]; }
    }
    // URDF rpy: R0 = Rz(yaw)·Ry(pitch)·Rx(roll)
    const double cr = std::cos(r[3]), sr = std::sin(r[3]), cpi = std::cos(r[4]), sp = std::sin(r[4]);
    const double cy = std::cos(r[5]), sy = std::sin(r[5]);
    double R0[9] = {cy * cpi, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                    sy * cpi, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                    -sp, cpi * sr, cpi * cr};
    if (r[3] == 0.0 && r[4] == 0.0 && r[5] == 0.0) { const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; std::memcpy(R0, I3, sizeof R0); }
    auto mm = [](const double* A, bool At, const double* Bm, double* out) {   // out = op(A)·B, 3×3 row-major
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          double acc = 0.0;
          for (int k = 0; k < 3; ++k) acc += (At ? A[3 * k + a] : A[3 * a + k]) * Bm[3 * k + b];
          out[3 * a + b] = acc;
        }
    };
    double T1[9], Rf[9];
    mm(Cprev, true, R0, T1); mm(T1, false, C, Rf);
    for (int k = 0; k < 9; ++k) c.Rf[i][k] = Rf[k];
    for (int a = 0; a < 3; ++a) {
      c.xyz[i][a] = Cprev[0 + a] * r[0] + Cprev[3 + a] * r[1] + Cprev[6 + a] * r[2];     // Cprevᵀ·xyz
      c.com[i][a] = C[0 + a] * r[10] + C[3 + a] * r[11] + C[6 + a] * r[12];              // Cᵀ·com
    }
    c.mass[i] = r[9];
    const double Il[9] = {r[13], r[14], r[15], r[14], r[16], r[17], r[15], r[17], r[18]};
    double T2[9], Ic[9];
    mm(C, true, Il, T2); mm(T2, false, C, Ic);                                           // Cᵀ·I·C
    c.I[i][0] = Ic[0]; c.I[i][1] = 0.5 * (Ic[1] + Ic[3]); c.I[i][2] = 0.5 * (Ic[2] + Ic[6]);
    c.I[i][3] = Ic[4]; c.I[i][4] = 0.5 * (Ic[5] + Ic[7]); c.I[i][5] = Ic[8];
    std::memcpy(Cprev, C, sizeof C);
    if (i == p.nq - 1)   // Izz + m (cx² + cy²) in the canonical frame
      c.last_diag = c.I[i][5] + c.mass[i] * (c.com[i][0] * c.com[i][0] + c.com[i][1] * c.com[i][1]);
  }
}

}  // namespace ilqr
