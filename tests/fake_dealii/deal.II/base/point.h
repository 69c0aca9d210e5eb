#pragma once
// stub (tests/fake_dealii/README.md): the repository's deal.II-free value types, re-exported as dealii::
#define MSFEM_SHIM_NAMESPACE dealii
#include "msfem/shim_types.hpp"
#undef MSFEM_SHIM_NAMESPACE
