// TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference `abawaca` driver.
//
// The reference hard-codes the SCG list path as a default member initialiser
// (/root/reference/src/abawaca.cpp:42) and has no flag for it.  Instead of
// patching a copy of the source, the reference translation unit is compiled
// where it lies (its own `main` stays the program's main) and a static
// initialiser, which runs before main, sets the public
// `Prog_params::scg_list_file` from $ABW_SCG_LIST.  `Prog_params::_instance`
// is constant-initialised to NULL (abawaca.cpp:49), so calling Instance() from
// a dynamic initialiser is well defined.  Nothing from the reference is copied
// into this repository.
#include <cstdlib>

#include "abawaca.cpp"  // resolved through -I/root/reference/src

namespace {
struct ScgListFromEnvironment {
	ScgListFromEnvironment()
	{
		const char* scg_list = getenv("ABW_SCG_LIST");
		if(scg_list != NULL)
			Prog_params::Instance()->scg_list_file = scg_list;
	}
} scg_list_from_environment;
}
