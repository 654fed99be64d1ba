#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <GU/GU_Detail.h>
class GU_PrimVDB : public GEO_PrimVDB {
   public:
	static GU_PrimVDB* buildFromGrid(GU_Detail& gdp, openvdb::GridBase::Ptr grid, const GEO_PrimVDB* src = nullptr, const char* name = nullptr);
};
