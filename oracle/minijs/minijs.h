// minijs -- a small JavaScript interpreter, TEST INFRASTRUCTURE ONLY.
//
// Purpose: execute the reference's own bundle, /root/reference/dist/Zlib-main.js, unmodified, so that the CPU oracle
// (oracle/zts_oracle.c) and the golden vectors under tests/golden/ are pinned to the real reference and not to a
// reading of it. The product (zlib.ts_b200/, bench.py's GPU arm) never links, loads or executes this.
//
// Language subset: everything the esbuild bundle uses -- var / const / let (function scoped), functions, arrows,
// classes (constructor, methods, static methods, getters / setters; no extends), object / array literals with
// shorthand properties and methods, default parameters, all ES5 operators plus ?? and **, if / for / while / do /
// switch / break / continue / return / throw / try-catch-finally, new, instanceof, typeof, void, in, delete.
// Semantics that matter for the hot path are the standard's: numbers are IEEE doubles, bitwise operators go through
// ToInt32 / ToUint32, typed arrays wrap on store and drop out-of-range writes, reads past the end give undefined,
// undefined + number = NaN, comparisons with NaN are false, arrays grow on write, `length` is live.
// Runtime: Uint8Array / Uint16Array / Uint32Array / Int8Array / Int16Array / Int32Array (+ ArrayBuffer sharing,
// set / subarray / slice / fill / map / forEach), Array, Object.assign / keys, Math, Number, String.fromCharCode,
// Date (UTC), Error, TextEncoder / TextDecoder, console.log, globalThis / window.
// Host functions for the harness scripts: readFile(path) -> Uint8Array, writeFile(path, bytes), print(...),
// scriptArgs (strings behind "--"), clock() (seconds).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mjs {

struct Interp;
Interp* create();
void destroy(Interp*);
// runs a script; returns false and fills `err` when it throws or does not parse
bool run_file(Interp*, const std::string& path, std::string& err);
bool run_source(Interp*, const std::string& src, const std::string& name, std::string& err);
void set_args(Interp*, const std::vector<std::string>& args);

}  // namespace mjs
