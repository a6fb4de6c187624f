// The reference's tests call std::sinf/std::cosf, which MSVC declares and libstdc++ 13 does not.
// Force-included (-include) when compiling the reference's unmodified test sources.
#include <cmath>
#include <math.h>
namespace std { using ::sinf; using ::cosf; }
