/* CCC/Gravity.h — the one physical constant the collection shares: standard gravity, as the value the reference
 * keeps in CCC::constants::g (reference include/CCC/Constants.h:10).  Same namespace and name, so code written
 * against the reference (`CCC::constants::g`) compiles unchanged. */
#pragma once
namespace CCC::constants
{
inline constexpr double g = 9.80665; // [m/s^2]
}
