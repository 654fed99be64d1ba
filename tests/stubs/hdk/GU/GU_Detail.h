#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <GEO/GEO_PrimVDB.h>
class GU_Detail {
   public:
	GA_Range getPrimitiveRange() const;
	const GEO_Primitive* getGEOPrimitive(GA_Offset off) const;
	GEO_Primitive* getGEOPrimitive(GA_Offset off);
	void clearAndDestroy();
};
