// chain_params.cuh — the POD description of a serial chain in canonical link frames (no CUDA dependencies: also read by
// the host-side builder, chain_host.hpp, and by the host compile of chain_lin.cuh in tests/).
#pragma once
#include <stdint.h>

namespace ilqr {

constexpr int kMaxQ = 8;

// Canonical link frames (built on the host from the URDF description, capi.cu): every link frame is
// re-oriented so that its joint axis is its own +z.  A joint is then "constant rotation Rf, then a turn
// about z" for every mechanism, which keeps the device code free of per-joint branches (the first version
// switched on the axis type per rotation and spent half its cycles on instruction-fetch stalls).
struct ChainP {
  int32_t nq;
  int32_t pad;
  double xyz[kMaxQ][3];       // joint origin in the (canonical) parent link frame
  double Rf[kMaxQ][9];        // row-major constant rotation: canonical child frame at q = 0 → canonical parent frame
  double mass[kMaxQ];
  double com[kMaxQ][3];       // in the canonical link frame
  double I[kMaxQ][6];         // ixx ixy ixz iyy iyz izz about the COM, canonical link axes
  double g[3];                // gravity acceleration in the base frame (fixed base only)
  double dt;
  // floating base (RBD_helper_functions.jl:7, floating = true): inertial of the root link, in its own frame
  double base_mass, base_com[3], base_I[6];
  // M[last][last]: inertia of the last link about its own joint axis (+z of its canonical frame) — independent of q
  double last_diag;
};

}  // namespace ilqr
