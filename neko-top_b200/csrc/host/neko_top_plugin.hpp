// C++17 host-side mirror of the reference plug-in types over the C ABI (include/neko_top_b200.h) -- what a
// compiled host (the reference's host code is compiled Fortran; no Fortran compiler exists in the build
// image) links against.  Header-only RAII wrappers, same names and argument order as the Fortran types:
//   advection_adjoint_t / adv_lin_no_dealias_t / adv_lin_dealias_t   adjoint/advection_adjoint.f90:43-82,
//                                                                    adjoint/adv_adjoint_no_dealias.f90:57-77,
//                                                                    adjoint/adv_adjoint_dealias.f90:56-131
//   simple_brinkman_source_term_t                                    source_terms/simple_brinkman_source_term.f90:52-153
// Device pointers are borrowed `void*` (Neko's field_t%x_d); errors follow the library's convention
// (abort by default, like neko_error) or throw std::runtime_error when abort-on-error is switched off.
#pragma once
#include <stdexcept>
#include <string>

#include "../../../include/neko_top_b200.h"

namespace neko_top_b200 {

inline void check(int status) {
  if (status != B200_OK) throw std::runtime_error(b200_last_error());
}

struct space_t { int lx; const double* dx; const double* wx; };                 // Xh%lx, Xh%dx, Xh%wx (host)
struct coef_t {                                                                  // device mirrors of coef_t
  space_t Xh; int nelv;
  const void *drdx, *dsdx, *dtdx, *drdy, *dsdy, *dtdy, *drdz, *dsdz, *dtdz, *B, *jacinv, *Binv;
};
struct field3 { void *x, *y, *z; };

class handle_t {
 public:
  handle_t(const coef_t& c, int device = 0, void* stream = nullptr) {
    check(b200_adjrhs_create(&h_, &c.Xh.lx, &c.nelv, &device));
    check(b200_adjrhs_set_stream(h_, stream));
    check(b200_adjrhs_set_space(h_, c.Xh.dx, c.Xh.wx));
    check(b200_adjrhs_set_geometry(h_, c.drdx, c.dsdx, c.dtdx, c.drdy, c.dsdy, c.dtdy, c.drdz, c.dsdz, c.dtdz, c.B));
  }
  handle_t(const handle_t&) = delete;
  handle_t& operator=(const handle_t&) = delete;
  ~handle_t() { b200_adjrhs_free(&h_); }
  void* get() const { return h_; }

 private:
  void* h_ = nullptr;
};

// advection_adjoint_t (abstract): compute_linear / compute_adjoint with f IN/OUT
class advection_adjoint_t {
 public:
  virtual ~advection_adjoint_t() = default;
  virtual void compute_linear(field3 v, field3 vb, field3 f) = 0;
  virtual void compute_adjoint(field3 v, field3 vb, field3 f) = 0;
};

class adv_lin_no_dealias_t : public advection_adjoint_t {
 public:
  explicit adv_lin_no_dealias_t(const coef_t& c) : c_(c), h_(c) {}
  void compute_adjoint(field3 v, field3 vb, field3 f) override {
    check(b200_adv_adjoint_compute(h_.get(), v.x, v.y, v.z, vb.x, vb.y, vb.z, f.x, f.y, f.z));
  }
  void compute_linear(field3 v, field3 vb, field3 f) override {
    check(b200_adv_linear_compute(h_.get(), v.x, v.y, v.z, vb.x, vb.y, vb.z, c_.jacinv, f.x, f.y, f.z));
  }
  handle_t& handle() { return h_; }

 protected:
  coef_t c_;
  handle_t h_;
};

class adv_lin_dealias_t : public advection_adjoint_t {
 public:
  // interp (lxd x lx), dxd (lxd x lxd), wd (lxd): GLL_to_GL and Xh_GL of init_dealias (:143-146), host, column-major
  adv_lin_dealias_t(int lxd, const coef_t& c, const double* interp, const double* dxd, const double* wd) : h_(c) {
    check(b200_adv_dealias_init(h_.get(), &lxd, interp, dxd, wd));
  }
  void compute_adjoint(field3 v, field3 vb, field3 f) override {
    check(b200_adv_adjoint_dealias_compute(h_.get(), v.x, v.y, v.z, vb.x, vb.y, vb.z, f.x, f.y, f.z));
  }
  void compute_linear(field3 v, field3 vb, field3 f) override {
    check(b200_adv_linear_dealias_compute(h_.get(), v.x, v.y, v.z, vb.x, vb.y, vb.z, f.x, f.y, f.z));
  }
  handle_t& handle() { return h_; }

 private:
  handle_t h_;
};

// simple_brinkman_source_term_t: init_from_components(f_x, f_y, f_z, design, u, v, w, coef) / compute_(t, tstep)
class simple_brinkman_source_term_t {
 public:
  void init_from_components(field3 f, const void* chi, field3 u, int n, void* stream = nullptr) {
    f_ = f; u_ = u; chi_ = chi; n_ = n; stream_ = stream;
  }
  void compute_(double /*t*/ = 0.0, int /*tstep*/ = 0) {
    check(b200_brinkman_compute(f_.x, f_.y, f_.z, u_.x, u_.y, u_.z, chi_, &n_, stream_));
  }

 private:
  field3 f_{}, u_{};
  const void* chi_ = nullptr;
  int n_ = 0;
  void* stream_ = nullptr;
};

// the fused path: adjoint_pnpn.f90:669-682 (+ :755-757 with step())
class fused_adjoint_rhs_t {
 public:
  explicit fused_adjoint_rhs_t(const coef_t& c, int device = 0, void* stream = nullptr) : h_(c, device, stream) {}
  void gs_init(const int64_t* key, bool on_device) {
    const int flag = on_device ? 1 : 0;
    check(b200_gs_init(h_.get(), key, &flag));
  }
  void compute(field3 v, field3 vb, const void* rho, field3 f, void* sens = nullptr) {
    check(b200_adjrhs_compute(h_.get(), v.x, v.y, v.z, vb.x, vb.y, vb.z, rho, nullptr, nullptr, nullptr, nullptr,
                              f.x, f.y, f.z, sens, nullptr));
  }
  void step(field3 v, field3 vb, const void* rho, field3 f, void* sens = nullptr) {
    check(b200_adjrhs_step(h_.get(), v.x, v.y, v.z, vb.x, vb.y, vb.z, rho, nullptr, nullptr, nullptr, nullptr,
                           f.x, f.y, f.z, sens, nullptr));
  }
  handle_t& handle() { return h_; }

 private:
  handle_t h_;
};

}  // namespace neko_top_b200
