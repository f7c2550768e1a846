// nimblesm_b200/host/view.h — non-owning strided host views with the interface of nimble::Viewify<N>
// (src/nimble_view.h:70-155).  The integrator's axpy updates (`v += a * w`, :181-214) run on the device in this
// build; the host views exist because callers of the ModelDataBase API (boundary conditions, output, tests)
// index nodal fields through them.
#pragma once
#include <array>
#include <cstddef>

namespace nimble_b200 {

template <std::size_t N = 2, class Scalar = double>
class Viewify
{
 public:
  Viewify() : data_(nullptr)
  {
    len_.fill(0);
    stride_.fill(0);
  }
  Viewify(Scalar* data, std::array<int, N> len, std::array<int, N> stride) : data_(data), len_(len), stride_(stride) {}
  template <std::size_t NN = N, typename = typename std::enable_if<(NN == 1)>::type>
  Viewify(Scalar* data, int len) : data_(data), len_({len}), stride_({1})
  {
  }
  template <std::size_t NN = N>
  typename std::enable_if<(NN == 1), Scalar>::type&
  operator()(int i) const
  {
    return data_[i];
  }
  template <std::size_t NN = N>
  typename std::enable_if<(NN == 2), Scalar>::type&
  operator()(int i, int j) const
  {
    return data_[i * stride_[0] + j * stride_[1]];
  }
  void
  zero()
  {
    const long n = (long)stride_[0] * len_[0];
    for (long i = 0; i < n; ++i) data_[i] = (Scalar)0;
  }
  void
  copy(const Viewify<N, Scalar>& ref)
  {
    const long n = (long)stride_[0] * len_[0];
    for (long i = 0; i < n; ++i) data_[i] = ref.data_[i];
  }
  Scalar*
  data() const
  {
    return data_;
  }
  std::array<int, N>
  size() const
  {
    return len_;
  }
  std::array<int, N>
  stride() const
  {
    return stride_;
  }

 protected:
  Scalar*            data_;
  std::array<int, N> len_;
  std::array<int, N> stride_;
};

}  // namespace nimble_b200
