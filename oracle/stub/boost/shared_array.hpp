// STUB (oracle/stub): boost::shared_array as far as include/stereo_binary_feature.h uses it (make_shared()).
#pragma once
#include <memory>
namespace boost {
template <typename T> class shared_array {
  std::shared_ptr<T> p_;
 public:
  explicit shared_array(T *p = nullptr) : p_(p, std::default_delete<T[]>()) {}
  T &operator[](std::ptrdiff_t i) const { return p_.get()[i]; }
  T *get() const { return p_.get(); }
};
}  // namespace boost
