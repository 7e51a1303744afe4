// ORACLE SHIM (test infrastructure): boost::shared_ptr as the reference uses it (construction from new, reset, ->, copy) maps onto std::shared_ptr.
#pragma once
#include <memory>
namespace boost {
template <typename T> using shared_ptr = std::shared_ptr<T>;
}
