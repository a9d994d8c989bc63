// Stand-in: boost::math::lgamma -> std::lgamma (agree to ~1e-15 relative).
#pragma once
#include <cmath>
namespace boost { namespace math { template <class T> inline double lgamma(T x) { return std::lgamma(static_cast<double>(x)); } } }
