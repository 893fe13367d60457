// Stub of OUR OWN (not reference code): glowl::BufferObject as far as the reference's CUDAQuickSurf.{h,cu} mention it (the GL vertex
// buffers of calc_surf's mesh hand-off, which the density harness never reaches).  glowl is a vcpkg dependency this image lacks.
#pragma once
#include <cstddef>
namespace glowl {
class BufferObject {
public:
    BufferObject(unsigned, const void*, std::size_t, unsigned) {}
    void rebuffer(const void*, std::size_t) {}
    unsigned getName() const { return 0; }
    std::size_t getByteSize() const { return 0; }
};
} // namespace glowl
