/* CCC/Contact.h — flat stand-in for ForceColl::Contact as the hot path reads it.
 *
 * The reference's DDP problems only read `contact->vertexWithRidgeList_` (vertex + friction-pyramid
 * ridges, reference src/DdpCentroidal.cpp:49-60) and `contact->ridgeNum()` (:21-30).  ForceColl is an
 * external dependency; this header restates that much of it without Eigen.
 */
#pragma once
#include <array>
#include <cmath>
#include <memory>
#include <string>
#include <vector>

namespace ForceColl
{
using Vector3d = std::array<double, 3>;

class Contact
{
public:
  struct VertexWithRidge
  {
    Vector3d vertex;
    std::vector<Vector3d> ridgeList;
  };

  virtual ~Contact() {}

  int ridgeNum() const
  {
    int n = 0;
    for(const auto & v : vertexWithRidgeList_) n += static_cast<int>(v.ridgeList.size());
    return n;
  }

  std::string name_;
  std::vector<VertexWithRidge> vertexWithRidgeList_;
};

/** Surface contact with identity orientation: each vertex carries the `ridge_num` unit ridges
 *  normalize(mu cos(2 pi i / n), mu sin(2 pi i / n), 1) of the linearised friction pyramid. */
class SurfaceContact : public Contact
{
public:
  SurfaceContact(const std::string & name, double fricCoeff, const std::vector<Vector3d> & localVertices, int ridge_num = 4)
  {
    name_ = name;
    std::vector<Vector3d> ridges;
    for(int i = 0; i < ridge_num; i++)
    {
      const double theta = 2 * 3.14159265358979323846 * (static_cast<double>(i) / ridge_num);
      Vector3d r = {fricCoeff * std::cos(theta), fricCoeff * std::sin(theta), 1.0};
      const double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      ridges.push_back({r[0] / n, r[1] / n, r[2] / n});
    }
    for(const auto & v : localVertices) vertexWithRidgeList_.push_back({v, ridges});
  }
};
} // namespace ForceColl
