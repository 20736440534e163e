/* CCC/detail/QpEngine.h — host-side owner of one batched QP structure on the CUDA engine (ccc_qp_*).
 *
 * Stands where the reference holds `std::shared_ptr<QpSolverCollection::QpSolver> qp_solver_` plus
 * `QpSolverCollection::QpCoeff qp_coeff_` (reference include/CCC/LinearMpcZmp.h:117-121,
 * include/CCC/IntrinsicallyStableMpc.h:113-117, include/CCC/LinearMpcXY.h:247-251): the matrices
 * (obj_mat_, eq_mat_, ineq_mat_) are fixed per structure and shared by the batch, the vectors
 * (obj_vec_, eq_vec_, ineq_vec_) carry a batch axis.  No CPU fallback: solve() throws without a GPU.
 */
#pragma once
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../../include/ccc_b200.h"
#include "Dense.h"

namespace QpSolverCollection
{
/** Kept for signature compatibility with the reference constructors; there is one backend here. */
enum class QpSolverType
{
  Any = -2,
  Uninitialized = -1,
  B200 = 100
};
} // namespace QpSolverCollection

namespace CCC
{
namespace detail
{
class QpEngine
{
public:
  QpEngine() {}
  ~QpEngine()
  {
    if(ws_) ccc_qp_destroy(ws_);
  }
  QpEngine(const QpEngine &) = delete;
  QpEngine & operator=(const QpEngine &) = delete;

  /** min 0.5 x'Qx + c'x  s.t.  A x = b, C x <= d   (Q n x n, A n_eq x n, C n_ineq x n). */
  void setup(const Matrix & Q, const Matrix & A, const Matrix & C)
  {
    n_ = Q.rows();
    n_eq_ = A.rows();
    n_ineq_ = C.rows();
    if(Q.cols() != n_ || (n_eq_ > 0 && A.cols() != n_) || C.cols() != n_) throw std::invalid_argument("QpEngine::setup: shapes");
    const bool same_dims = ws_ && Q.rows() == Q_.rows() && A.rows() == A_.rows() && C.rows() == C_.rows();
    Q_ = Q;
    A_ = A;
    C_ = C;
    // LinearMpcXY / LinearMpcZ call setup() every control tick with new matrices of (usually) the same shape: keep the
    // device workspace then and only re-upload / re-factorise (the reference re-runs QpCoeff::setup only when the
    // dimensions change, src/LinearMpcXY.cpp:134-138)
    if(!same_dims)
    {
      if(ws_) ccc_qp_destroy(ws_);
      ws_ = nullptr;
      ws_batch_ = 0;
    }
    matrices_resident_ = false;
  }

  int dimVar() const { return n_; }
  int dimEq() const { return n_eq_; }
  int dimIneq() const { return n_ineq_; }
  int batch() const { return batch_; }

  /** Size the per-problem vectors for a batch of B problems (c is all zeros unless with_obj_vec). */
  void resize(int B, bool with_obj_vec)
  {
    batch_ = B;
    with_c_ = with_obj_vec;
    c_.assign(with_obj_vec ? static_cast<size_t>(B) * n_ : 0, 0.0);
    b_.assign(static_cast<size_t>(B) * n_eq_, 0.0);
    d_.assign(static_cast<size_t>(B) * n_ineq_, 0.0);
  }
  double * objVec(int b) { return c_.data() + static_cast<size_t>(b) * n_; }
  double * eqVec(int b) { return b_.data() + static_cast<size_t>(b) * n_eq_; }
  double * ineqVec(int b) { return d_.data() + static_cast<size_t>(b) * n_ineq_; }

  /** Solve the batch; returns x [B][n].  Throws if the call itself fails.  A problem that is not solved (infeasible,
   *  iteration limit, active set full, Q not positive definite: status 1..4) is reported the way
   *  QpSolverCollection::QpSolver::solve does it — an error line on std::cerr and the last iterate returned —
   *  and counted in numFailed(); callers that must not act on such a result check status(b). */
  const std::vector<double> & solve()
  {
    if(!ws_ || batch_ > ws_batch_)
    {
      if(ws_) ccc_qp_destroy(ws_);
      ws_ = ccc_qp_create(n_, n_eq_, n_ineq_, batch_);
      if(!ws_) throw std::runtime_error(std::string("ccc_qp_create: ") + ccc_last_error());
      ws_batch_ = batch_;
      matrices_resident_ = false;
    }
    x_.assign(static_cast<size_t>(batch_) * n_, 0.0);
    iters_.assign(batch_, 0);
    status_.assign(batch_, 0);
    n_active_.assign(batch_, 0);
    active_.assign(static_cast<size_t>(batch_) * n_, -1);
    ccc_qp_batch_t bt{};
    bt.n = n_;
    bt.n_eq = n_eq_;
    bt.n_ineq = n_ineq_;
    bt.batch = batch_;
    // the matrices are uploaded and factorised by the first solve after setup(); later solves reuse them
    bt.Q = matrices_resident_ ? nullptr : Q_.data();
    bt.A = (matrices_resident_ || !n_eq_) ? nullptr : A_.data();
    bt.C = matrices_resident_ ? nullptr : C_.data();
    bt.c = with_c_ ? c_.data() : nullptr;
    bt.b = n_eq_ ? b_.data() : nullptr;
    bt.d = d_.data();
    ccc_qp_result_t rs{};
    rs.x = x_.data();
    rs.iters = iters_.data();
    rs.status = status_.data();
    rs.n_active = n_active_.data();
    rs.active = active_.data();
    const int rc = ccc_qp_solve(ws_, &bt, &rs, CCC_MEM_HOST, nullptr);
    if(rc != CCC_OK) throw std::runtime_error(std::string("ccc_qp_solve: ") + ccc_last_error());
    matrices_resident_ = true;
    n_failed_ = 0;
    int first = -1;
    for(int b = 0; b < batch_; b++)
      if(status_[b] != 0)
      {
        if(first < 0) first = b;
        n_failed_++;
      }
    if(n_failed_ > 0)
      std::cerr << "[CCC::QpEngine] failed to solve " << n_failed_ << " of " << batch_ << " QPs (first: problem " << first
                << ", status " << status_[first] << ": 1 infeasible, 2 iteration limit, 3 not positive definite)"
                << std::endl;
    return x_;
  }
  /** Problems of the last solve() whose status is not 0. */
  int numFailed() const { return n_failed_; }

  const double * x(int b) const { return x_.data() + static_cast<size_t>(b) * n_; }
  int status(int b) const { return status_[b]; }
  int iters(int b) const { return iters_[b]; }
  int numActive(int b) const { return n_active_[b]; }

private:
  int n_ = 0, n_eq_ = 0, n_ineq_ = 0, batch_ = 0, ws_batch_ = 0, n_failed_ = 0;
  bool with_c_ = false, matrices_resident_ = false;
  Matrix Q_, A_, C_;
  std::vector<double> c_, b_, d_, x_;
  std::vector<int32_t> iters_, status_, n_active_, active_;
  ccc_qp_ws_t * ws_ = nullptr;
};
} // namespace detail
} // namespace CCC
