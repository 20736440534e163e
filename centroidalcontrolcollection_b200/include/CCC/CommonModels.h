/* CCC/CommonModels.h — reference include/CCC/CommonModels.h, src/CommonModels.cpp:8-17.
 * ComZmpModelJerkInput: state (CoM position, velocity, acceleration), input CoM jerk, output ZMP. */
#pragma once
#include "Gravity.h"
#include "StateSpaceModel.h"

namespace CCC
{
class ComZmpModelJerkInput : public StateSpaceModel
{
public:
  explicit ComZmpModelJerkInput(double com_height) : StateSpaceModel(3, 1, 1)
  {
    A_(0, 1) = 1;
    A_(1, 2) = 1;
    B_(2, 0) = 1;
    C_(0, 0) = 1;
    C_(0, 2) = -1 * com_height / constants::g;
  }
};
} // namespace CCC
