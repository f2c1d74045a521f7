#include "launch_march.inl"
namespace odinn {
ODINN_INSTANTIATE_MARCH(double)
}
