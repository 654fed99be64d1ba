// TEST INFRASTRUCTURE — see ../openvdb.h.  openvdb::tree::LeafManager: a linear array of the tree's leaf nodes.
#pragma once

#include <type_traits>

#include "../openvdb.h"

namespace openvdb {
namespace tree {
template <typename TreeT>
class LeafManager {
   public:
	using LeafType = typename std::conditional<std::is_const<TreeT>::value, const typename TreeT::LeafNodeType, typename TreeT::LeafNodeType>::type;
	explicit LeafManager(TreeT& tree) {
		for (const auto& kv : tree.leafMap()) leaves_.push_back(kv.second.get());
	}
	size_t leafCount() const { return leaves_.size(); }
	LeafType& leaf(size_t i) const { return *leaves_[i]; }

   private:
	std::vector<LeafType*> leaves_;
};
}  // namespace tree
}  // namespace openvdb
