/* CCC/EigenInterop.h — conversions between the drop-in headers' vector types and Eigen's, for callers that hold the
 * reference's types (Eigen::Vector3d pos, Eigen::VectorXd planned_force_scales, ...).
 *
 * The drop-in classes of this directory use std::array<double, N> where the reference uses Eigen::Matrix<double, N, 1>
 * and std::vector<double> where it uses Eigen::VectorXd, because Eigen is not part of the image they are built and tested
 * in.  Where <Eigen/Core> is on the include path this header is active (CCC_B200_HAS_EIGEN == 1) and gives
 *
 *   CCC::toEigen(a)     std::array<double, N> -> Eigen::Matrix<double, N, 1>;  std::vector<double> -> Eigen::VectorXd
 *   CCC::toArray<N>(v)  any Eigen vector expression of N entries -> std::array<double, N>
 *   CCC::toVector(v)    any Eigen vector expression -> std::vector<double>
 *   CCC::toVectorList(l), CCC::toEigenList(l)   std::vector<Eigen::VectorXd> <-> std::vector<std::vector<double>> (u_list)
 *
 * so that a reference call site changes from
 *     initial_param.pos = sim.state_.pos.linear();                       Eigen::VectorXd u = ddp.planOnce(...);
 * to
 *     initial_param.pos = CCC::toArray<3>(sim.state_.pos.linear());      Eigen::VectorXd u = CCC::toEigen(ddp.planOnce(...));
 *
 * Without Eigen the header is empty.  It is compile-tested against a minimal stand-in for <Eigen/Core>
 * (tests/cpp/eigen_standin, tests/cpp/TestEigenInterop.cpp): only the Eigen members named here are used — Matrix<double, N, 1>
 * and VectorXd construction, size(), operator[], MatrixBase<Derived>, derived(), RowsAtCompileTime.
 */
#pragma once
#include <array>
#include <cstddef>
#include <stdexcept>
#include <vector>

#if !defined(CCC_B200_NO_EIGEN) && defined(__has_include)
#  if __has_include(<Eigen/Core>)
#    include <Eigen/Core>
#    define CCC_B200_HAS_EIGEN 1
#  endif
#endif
#ifndef CCC_B200_HAS_EIGEN
#  define CCC_B200_HAS_EIGEN 0
#endif

#if CCC_B200_HAS_EIGEN
namespace CCC
{
template<std::size_t N>
inline Eigen::Matrix<double, static_cast<int>(N), 1> toEigen(const std::array<double, N> & a)
{
  Eigen::Matrix<double, static_cast<int>(N), 1> v;
  for(std::size_t i = 0; i < N; i++) v[static_cast<int>(i)] = a[i];
  return v;
}

inline Eigen::VectorXd toEigen(const std::vector<double> & a)
{
  Eigen::VectorXd v(static_cast<int>(a.size()));
  for(std::size_t i = 0; i < a.size(); i++) v[static_cast<int>(i)] = a[i];
  return v;
}

/** Any Eigen vector expression with N entries (checked at run time for dynamic sizes). */
template<std::size_t N, class Derived>
inline std::array<double, N> toArray(const Eigen::MatrixBase<Derived> & m)
{
  const Eigen::Matrix<double, Derived::RowsAtCompileTime, 1> v = m.derived(); // evaluates expressions (a + b, block(), ...)
  if(static_cast<std::size_t>(v.size()) != N) throw std::invalid_argument("CCC::toArray: vector size differs from N");
  std::array<double, N> a;
  for(std::size_t i = 0; i < N; i++) a[i] = v[static_cast<int>(i)];
  return a;
}

template<class Derived>
inline std::vector<double> toVector(const Eigen::MatrixBase<Derived> & m)
{
  const Eigen::Matrix<double, Derived::RowsAtCompileTime, 1> v = m.derived();
  std::vector<double> a(static_cast<std::size_t>(v.size()));
  for(std::size_t i = 0; i < a.size(); i++) a[i] = v[static_cast<int>(i)];
  return a;
}

/** InitialParam::u_list of the DDP classes: std::vector<Eigen::VectorXd> in the reference. */
inline std::vector<std::vector<double>> toVectorList(const std::vector<Eigen::VectorXd> & l)
{
  std::vector<std::vector<double>> out;
  out.reserve(l.size());
  for(const auto & v : l) out.push_back(toVector(v));
  return out;
}

inline std::vector<Eigen::VectorXd> toEigenList(const std::vector<std::vector<double>> & l)
{
  std::vector<Eigen::VectorXd> out;
  out.reserve(l.size());
  for(const auto & v : l) out.push_back(toEigen(v));
  return out;
}
} // namespace CCC
#endif
