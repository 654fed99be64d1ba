#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <SYS/SYS_Types.h>
using GA_Offset = exint;
class GA_Range {};
class GA_Iterator {
   public:
	explicit GA_Iterator(const GA_Range& range);
	bool atEnd() const;
	GA_Iterator& operator++();
	GA_Offset operator*() const;
};
class GA_PrimitiveTypeId {
   public:
	bool operator==(const GA_PrimitiveTypeId& o) const;
	bool operator!=(const GA_PrimitiveTypeId& o) const;
};
