// Minimal stand-in for chrono/core/ChTypes.h: chrono_types::make_shared / make_unique (no Eigen alignment concerns here).
#ifndef CHRONO_B200_CHTYPES_H
#define CHRONO_B200_CHTYPES_H
#include <memory>
#include <utility>
namespace chrono_types {
template <typename T, typename... Args>
std::shared_ptr<T> make_shared(Args&&... args) { return std::make_shared<T>(std::forward<Args>(args)...); }
template <typename T, typename... Args>
std::unique_ptr<T> make_unique(Args&&... args) { return std::make_unique<T>(std::forward<Args>(args)...); }
}  // namespace chrono_types
#endif
