// TEST INFRASTRUCTURE: replacement for the reference's transform_output_iterator.h when building
// oracle/_ref with the toolkit's CUB (CCCL 2.8), whose DeviceSelect needs a random-access output
// iterator (operator+ / operator+=); the reference pins CCCL 2.3.2 where ++ was enough.  Same class
// name and constructor, written for this build only.
#ifndef ORC_SHIM_TRANSFORM_OUTPUT_ITERATOR_H_
#define ORC_SHIM_TRANSFORM_OUTPUT_ITERATOR_H_
#include <cstddef>
#include <iterator>
namespace frc971::apriltag {
template <typename InputType, typename OutputType, typename ConversionOp, typename OffsetT = ptrdiff_t>
class TransformOutputIterator {
  struct Proxy {
    OutputType *ptr;
    ConversionOp op;
    __host__ __device__ Proxy(OutputType *p, ConversionOp o) : ptr(p), op(o) {}
    __host__ __device__ Proxy operator=(InputType val) {
      *ptr = op(val);
      return *this;
    }
  };

 public:
  using self_type = TransformOutputIterator;
  using difference_type = OffsetT;
  using value_type = void;
  using pointer = void *;
  using reference = Proxy;
  using iterator_category = std::random_access_iterator_tag;

  __host__ __device__ TransformOutputIterator(OutputType *ptr, const ConversionOp op) : op_(op), ptr_(ptr) {}
  __host__ __device__ self_type operator++(int) { self_type r = *this; ptr_++; return r; }
  __host__ __device__ self_type &operator++() { ptr_++; return *this; }
  __host__ __device__ self_type operator--(int) { self_type r = *this; ptr_--; return r; }
  __host__ __device__ self_type &operator--() { ptr_--; return *this; }
  template <typename D> __host__ __device__ self_type operator+(D n) const { return self_type(ptr_ + n, op_); }
  template <typename D> __host__ __device__ self_type operator-(D n) const { return self_type(ptr_ - n, op_); }
  template <typename D> __host__ __device__ self_type &operator+=(D n) { ptr_ += n; return *this; }
  template <typename D> __host__ __device__ self_type &operator-=(D n) { ptr_ -= n; return *this; }
  __host__ __device__ difference_type operator-(const self_type &o) const { return ptr_ - o.ptr_; }
  __host__ __device__ reference operator*() const { return Proxy(ptr_, op_); }
  template <typename D> __host__ __device__ reference operator[](D n) const { return Proxy(ptr_ + n, op_); }
  __host__ __device__ bool operator==(const self_type &rhs) const { return ptr_ == rhs.ptr_; }
  __host__ __device__ bool operator!=(const self_type &rhs) const { return ptr_ != rhs.ptr_; }

 private:
  ConversionOp op_;
  OutputType *ptr_;
};
}  // namespace frc971::apriltag
#endif
