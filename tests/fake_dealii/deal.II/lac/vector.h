#pragma once
#include <deal.II/base/point.h>
