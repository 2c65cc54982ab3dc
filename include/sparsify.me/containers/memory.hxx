// containers/memory.hxx -- where a container's storage lives.
// Same enumerators as the reference (include/sparsify.me/containers/memory.hxx:13): the
// drivers spell them `memory_space_t::device` / `memory_space_t::host`.
#pragma once
namespace sparsifyme {
enum memory_space_t { device, host };
}  // namespace sparsifyme
