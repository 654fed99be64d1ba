// Host-side fp32 -> fp16 conversions for the split-fp16 weight streams (encode_tc_host.cpp, encode_tc128_host.cpp):
// w ~= hi + lo / 2048 with hi = fp16(w), lo = fp16((w - hi) * 2048).
#pragma once

#include <cstdint>
#include <cstring>

namespace vqvdb {

// IEEE binary32 -> binary16, round to nearest even, subnormals kept.
inline uint16_t f32_to_f16_rn(float f) {
	uint32_t x;
	std::memcpy(&x, &f, 4);
	const uint32_t sign = (x >> 16) & 0x8000u;
	const uint32_t abs = x & 0x7fffffffu;
	if (abs >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (abs > 0x7f800000u ? 0x200u : 0u));
	if (abs >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);  // >= 65520 rounds to infinity
	if (abs <= 0x33000000u) return (uint16_t)sign;              // <= 2^-25 rounds to zero (tie goes to even)
	const int exp = (int)(abs >> 23) - 127;
	const uint32_t mant = (abs & 0x7fffffu) | 0x800000u;
	if (exp < -14) {  // subnormal result: units of 2^-24
		const int shift = (-14 - exp) + 13;
		uint32_t q = mant >> shift;
		const uint32_t rem = mant & ((1u << shift) - 1u), half = 1u << (shift - 1);
		if (rem > half || (rem == half && (q & 1u))) ++q;
		return (uint16_t)(sign | q);
	}
	uint32_t q = ((uint32_t)(exp + 15) << 10) | ((mant & 0x7fffffu) >> 13);
	const uint32_t rem = mant & 0x1fffu;
	if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) ++q;  // a carry moves into the exponent field correctly
	return (uint16_t)(sign | q);
}

inline float f16_to_f32(uint16_t h) {
	const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 0x1fu, mant = h & 0x3ffu, out;
	if (exp == 0) {
		if (mant == 0) out = sign;
		else {
			int e = -1;
			do {
				mant <<= 1;
				++e;
			} while (!(mant & 0x400u));
			out = sign | ((uint32_t)(127 - 15 - e) << 23) | ((mant & 0x3ffu) << 13);
		}
	} else if (exp == 31) out = sign | 0x7f800000u | (mant << 13);
	else out = sign | ((exp - 15 + 127) << 23) | (mant << 13);
	float f;
	std::memcpy(&f, &out, 4);
	return f;
}

// part 0: fp16(w); part 1: fp16((w - fp16(w)) * 2048)
inline uint16_t split_part(float w, int part) {
	const uint16_t hi = f32_to_f16_rn(w);
	if (part == 0) return hi;
	return f32_to_f16_rn((w - f16_to_f32(hi)) * 2048.f);
}

}  // namespace vqvdb
