// TEST INFRASTRUCTURE — stand-in for <openvdb/Types.h>, only as much of it as the reference's .vqvdb container code
// (/root/reference/src/Utils/VQVDB_Reader.{hpp,cpp}) touches: openvdb::Coord as three int32 (sizeof == 12, which is
// what the record layout depends on: VQVDB_Reader.cpp:108,146-147) and openvdb::math::Mat4s as 16 floats with
// identity(), asPointer() and construction from a float array (VQVDB_Reader.hpp:26,29; VQVDB_Reader.cpp:122,210).
// With it the reference's reader/writer compile UNMODIFIED into oracle/_ref/libvqvdb_fmt.so (oracle/Makefile: fmt),
// the format oracle of tests/test_vqvdb_file.py.  OpenVDB itself is not available in this image.
#pragma once

#include <cstdint>
#include <cstring>

namespace openvdb {
namespace math {
class Coord {
   public:
	Coord() : v_{0, 0, 0} {}
	Coord(int32_t x, int32_t y, int32_t z) : v_{x, y, z} {}
	int32_t x() const { return v_[0]; }
	int32_t y() const { return v_[1]; }
	int32_t z() const { return v_[2]; }

   private:
	int32_t v_[3];
};
class Mat4s {
   public:
	Mat4s() { identity(); }
	explicit Mat4s(const float* a) { std::memcpy(m_, a, sizeof(m_)); }
	void identity() {
		for (int i = 0; i < 16; ++i) m_[i] = (i % 5 == 0) ? 1.0f : 0.0f;
	}
	float* asPointer() { return m_; }
	const float* asPointer() const { return m_; }

   private:
	float m_[16];
};
}  // namespace math
using Coord = math::Coord;
}  // namespace openvdb
static_assert(sizeof(openvdb::Coord) == 12, "openvdb::Coord is three int32");
