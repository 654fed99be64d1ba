// TEST INFRASTRUCTURE — compiles and RUNS vqvdb_b200/cpp/openvdb_adapter.hpp against the functional OpenVDB stand-in of
// tests/stubs/openvdb/ (no OpenVDB in this image), through the repository's own orchestrator (VQVAECodec::compress /
// decompress, .vqvdb v3 container) with a deterministic CPU stand-in for the backend: grid -> leaves -> file -> leaves ->
// grid.  What it pins: leaf order and origins, the 512-float leaf-buffer layout, the 16-float transform, grid names,
// setValuesOn on every decoded leaf, empty grids skipped (VQVAECodec.cpp:89-92), multi-grid files, ragged batch sizes.
#define VQVDB_B200_WITH_OPENVDB 1
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "openvdb_adapter.hpp"

namespace {
// index j of a leaf = round(255 * mean of the 8 voxels of 2x2x2 cell j); decode = that value / 255 on the whole cell
class CellMeanBackend final : public IVQVAECodec {
   public:
	mutable int encodeCalls = 0, decodeCalls = 0;
	Tensor encode(const TensorView& v) const override {
		if (v.dtype != DataType::FLOAT32) throw std::runtime_error("encode expects FLOAT32 data.");
		++encodeCalls;
		const int64_t n = v.shape.at(0);
		Tensor out;
		out.dtype = DataType::UINT8;
		out.shape = {n, 4, 4, 4};
		out.buffer.resize((size_t)n * 64);
		const float* x = static_cast<const float*>(v.data);
		for (int64_t l = 0; l < n; ++l)
			for (int c = 0; c < 64; ++c) {
				const int cx = c >> 4, cy = (c >> 2) & 3, cz = c & 3;
				float s = 0.f;
				for (int d = 0; d < 8; ++d) s += x[l * 512 + (((2 * cx + (d >> 2)) << 6) | ((2 * cy + ((d >> 1) & 1)) << 3) | (2 * cz + (d & 1)))];
				out.getData<uint8_t>()[l * 64 + c] = (uint8_t)std::lround(255.f * s / 8.f);
			}
		return out;
	}
	Tensor decode(const TensorView& v) const override {
		if (v.dtype != DataType::UINT8) throw std::runtime_error("decode expects UINT8 data.");
		++decodeCalls;
		const int64_t n = v.shape.at(0);
		Tensor out;
		out.dtype = DataType::FLOAT32;
		out.shape = {n, 1, 8, 8, 8};
		out.buffer.resize((size_t)n * 512 * sizeof(float));
		const uint8_t* idx = static_cast<const uint8_t*>(v.data);
		for (int64_t l = 0; l < n; ++l)
			for (int o = 0; o < 512; ++o) {
				const int x = o >> 6, y = (o >> 3) & 7, z = o & 7;
				out.getData<float>()[l * 512 + o] = idx[l * 64 + (((x >> 1) << 4) | ((y >> 1) << 2) | (z >> 1))] / 255.f;
			}
		return out;
	}
	const std::vector<int64_t>& getLatentShape() const override { return shape_; }

   private:
	std::vector<int64_t> shape_{4, 4, 4};
};

#define CHECK(cond)                                                              \
	do {                                                                         \
		if (!(cond)) {                                                           \
			std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
			std::exit(1);                                                        \
		}                                                                        \
	} while (0)

openvdb::FloatGrid::Ptr makeGrid(const std::string& name, int nLeaves, double scale, unsigned seed) {
	auto grid = openvdb::FloatGrid::create(0.0f);
	grid->setName(name);
	openvdb::Mat4R m;
	for (int i = 0; i < 3; ++i) m.asPointer()[i * 5] = scale;
	m.asPointer()[12] = 1.5; m.asPointer()[13] = -2.0; m.asPointer()[14] = 0.25;
	grid->setTransform(openvdb::math::Transform::createLinearTransform(m));
	auto acc = grid->getAccessor();
	unsigned s = seed;
	auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) & 0xffff; };
	for (int l = 0; l < nLeaves; ++l) {
		auto* leaf = acc.touchLeaf(openvdb::Coord(8 * (int)(rnd() % 64) - 256, 8 * (int)(rnd() % 64) - 256, 8 * l));
		for (int o = 0; o < 512; ++o) leaf->buffer().data()[o] = (float)(rnd() % 256) / 255.f;
		leaf->setValuesOn();
	}
	return grid;
}
}  // namespace

int main(int argc, char** argv) {
	const std::string path = argc > 1 ? argv[1] : "/tmp/adapter_stub_test.vqvdb";
	auto g0 = makeGrid("density", 150, 0.1, 1), g1 = makeGrid("empty", 0, 1.0, 2), g2 = makeGrid("temperature", 7, 0.5, 3);

	// grid -> leaves: LeafManager order, origins, buffers, transform
	const LeafGrid flat = vqvdb_openvdb::toLeafGrid(*g0);
	CHECK(flat.name == "density" && flat.leafCount() == 150 && flat.voxels.size() == 150 * 512);
	CHECK(std::fabs(flat.transform[0] - 0.1f) < 1e-7f && flat.transform[12] == 1.5f && flat.transform[13] == -2.0f && flat.transform[15] == 1.0f);
	{
		const openvdb::tree::LeafManager<const openvdb::FloatTree> lm(g0->tree());
		for (size_t i = 0; i < lm.leafCount(); ++i) {
			CHECK(flat.origins[i].x == lm.leaf(i).origin().x() && flat.origins[i].y == lm.leaf(i).origin().y() && flat.origins[i].z == lm.leaf(i).origin().z());
			CHECK(std::memcmp(flat.voxels.data() + i * 512, lm.leaf(i).buffer().data(), 2048) == 0);
		}
	}
	// leaves -> grid: exact inverse, every voxel active
	{
		const auto back = vqvdb_openvdb::toFloatGrid(flat);
		CHECK(back->getName() == "density" && back->tree().leafCount() == 150);
		const LeafGrid again = vqvdb_openvdb::toLeafGrid(*back);
		CHECK(again.voxels == flat.voxels);
		for (const auto& kv : back->tree().leafMap()) CHECK(kv.second->onVoxelCount() == 512);
		for (int i = 0; i < 16; ++i) CHECK(again.transform[i] == flat.transform[i]);
	}
	// the reference's entry points on grids, through the orchestrator and the .vqvdb container
	auto backend = std::make_unique<CellMeanBackend>();
	const CellMeanBackend* probe = backend.get();
	const VQVAECodec codec(std::move(backend));
	vqvdb_openvdb::compress(codec, {g0, g1, g2}, path);
	CHECK(probe->encodeCalls == 2);                       // one call per non-empty grid: the whole grid at once
	const auto out = vqvdb_openvdb::decompress(codec, path);
	CHECK(out.size() == 2 && out[0]->getName() == "density" && out[1]->getName() == "temperature");
	CHECK(out[0]->tree().leafCount() == 150 && out[1]->tree().leafCount() == 7);
	const LeafGrid want0 = vqvdb_openvdb::toLeafGrid(*g0), got0 = vqvdb_openvdb::toLeafGrid(*out[0]);
	for (int i = 0; i < 16; ++i) CHECK(got0.transform[i] == want0.transform[i]);
	for (size_t l = 0; l < 150; ++l) {
		CHECK(got0.origins[l].x == want0.origins[l].x && got0.origins[l].y == want0.origins[l].y && got0.origins[l].z == want0.origins[l].z);
		for (int c = 0; c < 64; ++c) {                     // decoded cell value == the rounded cell mean of the input leaf
			const int cx = c >> 4, cy = (c >> 2) & 3, cz = c & 3;
			float s = 0.f;
			for (int d = 0; d < 8; ++d) s += want0.voxels[l * 512 + (((2 * cx + (d >> 2)) << 6) | ((2 * cy + ((d >> 1) & 1)) << 3) | (2 * cz + (d & 1)))];
			const float expect = (float)std::lround(255.f * s / 8.f) / 255.f;
			CHECK(got0.voxels[l * 512 + ((2 * cx) << 6 | (2 * cy) << 3 | (2 * cz))] == expect);
			CHECK(got0.voxels[l * 512 + ((2 * cx + 1) << 6 | (2 * cy + 1) << 3 | (2 * cz + 1))] == expect);
		}
	}
	// the SOPs' batch sizes (64 default; ragged last batch) give the same file
	{
		std::vector<LeafGrid> flatAll{vqvdb_openvdb::toLeafGrid(*g0), vqvdb_openvdb::toLeafGrid(*g2)};
		codec.compress(flatAll, path + ".b64", 64);
		std::vector<LeafGrid> a, b;
		codec.decompress(path + ".b64", a, 64);
		codec.decompress(path, b, 0);
		CHECK(a.size() == 2 && b.size() == 2 && a[0].voxels == b[0].voxels && a[1].voxels == b[1].voxels);
		std::remove((path + ".b64").c_str());
	}
	std::remove(path.c_str());
	std::puts("adapter_stub_test: ok");
	return 0;
}
