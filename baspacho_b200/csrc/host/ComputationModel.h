// Polynomial timing models of the four dense op families of the factorization, used by the
// supernode-merge heuristic. Functional forms follow reference baspacho/baspacho/ComputationModel.h:57-99
// (potrf cubic in n; trsm quadratic-in-n times affine-in-k; syrk/gemm symmetric in (m,n), affine in k;
// asmbl bilinear in block rows/cols). Eigen-free: coefficients are plain arrays.
#pragma once
#include <array>

namespace BaSpaCho {

struct Lin2 {  // a + b*x
  double a = 0, b = 0;
  Lin2& operator+=(const Lin2& o) { a += o.a; b += o.b; return *this; }
  Lin2& operator-=(const Lin2& o) { a -= o.a; b -= o.b; return *this; }
  double operator[](int i) const { return i == 0 ? a : b; }
};

struct ComputationModel {
  std::array<double, 4> potrfParams{};
  std::array<double, 6> trsmParams{};
  std::array<double, 6> sygeParams{};
  std::array<double, 4> asmblParams{};

  double potrfEst(double n) const {
    const auto& p = potrfParams;
    return p[0] + n * (p[1] + n * (p[2] + n * p[3]));
  }
  double trsmEst(double n, double k) const {
    const auto& p = trsmParams;
    return p[0] + n * (p[1] + n * p[2]) + k * (p[3] + n * (p[4] + n * p[5]));
  }
  double sygeEst(double m, double n, double k) const {
    Lin2 l = sygeLinEst(m, n);
    return l.a + k * l.b;
  }
  double asmblEst(double br, double bc) const {
    Lin2 l = asmblLinEst(br);
    return l.a + bc * l.b;
  }
  // syrk/gemm time as an affine function of k (the source node width)
  Lin2 sygeLinEst(double m, double n) const {
    const auto& p = sygeParams;
    double u = m + n, v = m * n;
    return {p[0] + u * p[1] + v * p[2], p[3] + u * p[4] + v * p[5]};
  }
  // assemble time as an affine function of the number of block columns
  Lin2 asmblLinEst(double br) const {
    const auto& p = asmblParams;
    return {p[0] + br * p[1], p[2] + br * p[3]};
  }

  // presets published by the reference (ComputationModel.cpp:12-30) and a B200 preset for this backend
  static const ComputationModel model_OpenBlas_i7_1185g7;
  static const ComputationModel model_Cuda117_2080Ti;
  static const ComputationModel model_B200;
};

}  // namespace BaSpaCho
