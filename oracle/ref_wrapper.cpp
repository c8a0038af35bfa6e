// Builds the REFERENCE's own CPU ops (mega_core/csrc/{vision.cpp,cpu/*.cpp}) from where they
// lie under /root/reference, unmodified, into oracle/_ref/ -- test infrastructure only.
//
// The reference sources predate torch 1.11: AT_DISPATCH_FLOATING_TYPES(x.type(), ...) hands a
// DeprecatedTypeProperties to ::detail::scalar_type(), for which current torch has no overload
// (ROIAlign_cpu.cpp:242, nms_cpu.cpp:71). Supplying that overload here lets the files compile
// verbatim; no reference source is copied or edited.
#include <torch/extension.h>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail

#include "cpu/nms_cpu.cpp"
#include "cpu/ROIAlign_cpu.cpp"
#include "vision.cpp"
