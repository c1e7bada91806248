// stand-in for bonsai's encoder.h (un-vendored; test infrastructure, see spumoni_ref_shim.hpp). Minimizer digestion is
// outside the path this repo reproduces (reads are queried with -n): the types exist so that the reference's spumoni.cpp
// and refbuilder.cpp compile; using them is an error.
#pragma once
#include <array>
#include <memory>
#include <stdexcept>

#include "spumoni_ref_shim.hpp"

namespace bns {
enum RollingHashingType { DNA, PROTEIN };
namespace score {
struct Lex {};
}  // namespace score
[[noreturn]] inline void no_digestion() { throw std::runtime_error("minimizer digestion is not available in oracle/_ref (bonsai is not vendored): use -n"); }
template <class T>
struct RollingHasher {
    template <class... A> RollingHasher(A&&...) {}
    template <class F> void for_each_uncanon(F&&, const char*, size_t) { no_digestion(); }
    template <class F> void for_each_hash(F&&, const char*, size_t) { no_digestion(); }
};
struct Spacer {
    template <class... A> Spacer(A&&...) {}
    std::string to_string(uint64_t) const { no_digestion(); }
};
template <class Score, class T>
struct Encoder {
    template <class... A> Encoder(A&&...) {}
    template <class F> void for_each(F&&, const char*, size_t) { no_digestion(); }
};
}  // namespace bns
