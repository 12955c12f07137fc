// TEST INFRASTRUCTURE (oracle) -- stands in for <exprtk.hpp> (un-vendored,
// reference MathCode.h:4).  The oracle only feeds numeric load tables, so the
// three class templates MathCode.h names are empty.
#pragma once
namespace exprtk {
template <typename T> struct symbol_table {};
template <typename T> struct expression {};
template <typename T> struct parser {};
}
