// TEST INFRASTRUCTURE (oracle): forwards to the OpenCV stand-in, see core.hpp.
#include "core.hpp"
