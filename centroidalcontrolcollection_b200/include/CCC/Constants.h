/* CCC/Constants.h — reference include/CCC/Constants.h:10 */
#pragma once
namespace CCC
{
namespace constants
{
//! Gravitational acceleration [m/s^2]
constexpr double g = 9.80665;
} // namespace constants
} // namespace CCC
