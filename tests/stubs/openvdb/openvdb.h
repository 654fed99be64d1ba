// TEST INFRASTRUCTURE — a small, FUNCTIONAL stand-in for the part of the OpenVDB public API that
// vqvdb_b200/cpp/openvdb_adapter.hpp uses (OpenVDB itself is not installed in this image): a FloatGrid whose tree is an
// ordered map of 8^3 leaves, a value accessor with touchLeaf, LeafManager, Coord, Mat4d and a linear Transform.
// Signatures follow openvdb 11 (openvdb/Grid.h, tree/LeafNode.h, tree/LeafManager.h, math/Maps.h, math/Transform.h),
// so the adapter that compiles and runs against this stub is written against the real call shapes.
#pragma once

#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace openvdb {

using Real = double;
using Index = uint32_t;

namespace math {
class Coord {
   public:
	Coord() : v_{0, 0, 0} {}
	Coord(int32_t x, int32_t y, int32_t z) : v_{x, y, z} {}
	int32_t x() const { return v_[0]; }
	int32_t y() const { return v_[1]; }
	int32_t z() const { return v_[2]; }
	bool operator<(const Coord& o) const { return v_ < o.v_; }
	bool operator==(const Coord& o) const { return v_ == o.v_; }

   private:
	std::array<int32_t, 3> v_;
};

template <typename T>
class Mat4 {
   public:
	Mat4() { setIdentity(); }
	void setIdentity() {
		for (int i = 0; i < 16; ++i) m_[i] = (i % 5 == 0) ? T(1) : T(0);
	}
	T* asPointer() { return m_; }
	const T* asPointer() const { return m_; }

   private:
	T m_[16];
};
using Mat4d = Mat4<double>;
using Mat4s = Mat4<float>;
using Mat4R = Mat4<Real>;

class AffineMap {
   public:
	using Ptr = std::shared_ptr<AffineMap>;
	using ConstPtr = std::shared_ptr<const AffineMap>;
	explicit AffineMap(const Mat4d& m) : m_(m) {}
	Mat4d getMat4() const { return m_; }

   private:
	Mat4d m_;
};

class MapBase {
   public:
	using Ptr = std::shared_ptr<MapBase>;
	using ConstPtr = std::shared_ptr<const MapBase>;
	explicit MapBase(const Mat4d& m) : affine_(std::make_shared<AffineMap>(m)) {}
	AffineMap::Ptr getAffineMap() const { return affine_; }

   private:
	AffineMap::Ptr affine_;
};

class Transform {
   public:
	using Ptr = std::shared_ptr<Transform>;
	Transform() : map_(std::make_shared<MapBase>(Mat4d())) {}
	explicit Transform(const Mat4R& m) : map_(std::make_shared<MapBase>(m)) {}
	static Ptr createLinearTransform(const Mat4R& m) { return std::make_shared<Transform>(m); }
	MapBase::ConstPtr baseMap() const { return map_; }

   private:
	MapBase::Ptr map_;
};
}  // namespace math

using Coord = math::Coord;
using Mat4d = math::Mat4d;
using Mat4R = math::Mat4R;
using Mat4s = math::Mat4s;

namespace tree {
template <typename T>
class LeafBuffer {
   public:
	T* data() { return v_; }
	const T* data() const { return v_; }

   private:
	T v_[512] = {};
};

template <typename T>
class LeafNode {
   public:
	static const Index SIZE = 512, DIM = 8;
	LeafNode(const Coord& origin, const T& background) : origin_(origin) {
		for (Index i = 0; i < SIZE; ++i) buffer_.data()[i] = background;
	}
	const Coord& origin() const { return origin_; }
	LeafBuffer<T>& buffer() { return buffer_; }
	const LeafBuffer<T>& buffer() const { return buffer_; }
	void setValuesOn() { on_ = SIZE; }
	void setValueOn(Index offset, const T& v) {  // offset = (x<<6)|(y<<3)|z, leaf-local
		buffer_.data()[offset] = v;
		if (on_ < SIZE) ++on_;
	}
	Index onVoxelCount() const { return on_; }

   private:
	Coord origin_;
	LeafBuffer<T> buffer_;
	Index on_ = 0;
};

template <typename T>
class Tree {
   public:
	using ValueType = T;
	using LeafNodeType = LeafNode<T>;
	explicit Tree(const T& background) : background_(background) {}
	Index leafCount() const { return (Index)leaves_.size(); }
	LeafNodeType* touchLeaf(const Coord& xyz) {
		const Coord o(xyz.x() & ~7, xyz.y() & ~7, xyz.z() & ~7);
		auto it = leaves_.find(o);
		if (it == leaves_.end()) it = leaves_.emplace(o, std::make_unique<LeafNodeType>(o, background_)).first;
		return it->second.get();
	}
	const std::map<Coord, std::unique_ptr<LeafNodeType>>& leafMap() const { return leaves_; }

   private:
	T background_;
	std::map<Coord, std::unique_ptr<LeafNodeType>> leaves_;
};

template <typename TreeT>
class ValueAccessor {
   public:
	explicit ValueAccessor(TreeT& t) : tree_(&t) {}
	typename TreeT::LeafNodeType* touchLeaf(const Coord& xyz) { return tree_->touchLeaf(xyz); }

   private:
	TreeT* tree_;
};
}  // namespace tree

using FloatTree = tree::Tree<float>;

class GridBase {
   public:
	using Ptr = std::shared_ptr<GridBase>;
	using ConstPtr = std::shared_ptr<const GridBase>;
	virtual ~GridBase() = default;
	std::string getName() const { return name_; }
	void setName(const std::string& n) { name_ = n; }

   private:
	std::string name_;
};

template <typename TreeT>
class Grid : public GridBase {
   public:
	using Ptr = std::shared_ptr<Grid>;
	using ConstPtr = std::shared_ptr<const Grid>;
	using TreeType = TreeT;
	using Accessor = tree::ValueAccessor<TreeT>;
	explicit Grid(const typename TreeT::ValueType& background) : tree_(background), xform_(std::make_shared<math::Transform>()) {}
	static Ptr create(const typename TreeT::ValueType& background = typename TreeT::ValueType()) { return std::make_shared<Grid>(background); }
	const math::Transform& transform() const { return *xform_; }
	void setTransform(math::Transform::Ptr t) { xform_ = std::move(t); }
	TreeT& tree() { return tree_; }
	const TreeT& tree() const { return tree_; }
	Accessor getAccessor() { return Accessor(tree_); }

   private:
	TreeT tree_;
	math::Transform::Ptr xform_;
};
using FloatGrid = Grid<FloatTree>;

template <typename GridType>
inline typename GridType::Ptr gridPtrCast(const GridBase::Ptr& grid) { return std::dynamic_pointer_cast<GridType>(grid); }
template <typename GridType>
inline typename GridType::ConstPtr gridConstPtrCast(const GridBase::ConstPtr& grid) { return std::dynamic_pointer_cast<const GridType>(grid); }

inline void initialize() {}

}  // namespace openvdb
