// TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference `abawaca` driver.
//
// The reference hard-codes the SCG list path as a default member initialiser
// (/root/reference/src/abawaca.cpp:42) and has no flag for it.  Instead of
// patching a copy of the source we compile the reference translation unit
// where it lies, with its `main` renamed, and set the public
// `Prog_params::scg_list_file` from $ABW_SCG_LIST before handing control to it.
// Nothing from the reference is copied into this repository.
#include <cstdlib>

#define main abawaca_reference_main
#include "abawaca.cpp"  // resolved through -I/root/reference/src
#undef main

int main(int argc, const char** argv)
{
	const char* scg_list = getenv("ABW_SCG_LIST");
	if(scg_list != NULL)
		Prog_params::Instance()->scg_list_file = scg_list;
	return abawaca_reference_main(argc, argv);
}
