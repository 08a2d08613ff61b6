// java_compat.hpp — the JDK behaviours the reference's hot path depends on,
// restated for a JVM-less host: String.compareTo / String.hashCode over UTF-16
// code units and the iteration order of a java.util.HashMap<String, ?>.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace lagassign_host {

// UTF-8 -> UTF-16 code units (lenient: malformed bytes map to themselves).
inline std::vector<uint16_t> utf16_units(const std::string& s) {
    std::vector<uint16_t> out;
    out.reserve(s.size());
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s.data());
    const size_t n = s.size();
    size_t i = 0;
    auto cont = [&](size_t k) { return k < n && (p[k] & 0xC0) == 0x80; };
    while (i < n) {
        uint32_t cp;
        unsigned char c = p[i];
        if (c < 0x80) { cp = c; i += 1; }
        else if ((c & 0xE0) == 0xC0 && cont(i + 1)) { cp = ((c & 0x1Fu) << 6) | (p[i + 1] & 0x3Fu); i += 2; }
        else if ((c & 0xF0) == 0xE0 && cont(i + 1) && cont(i + 2)) {
            cp = ((c & 0x0Fu) << 12) | ((p[i + 1] & 0x3Fu) << 6) | (p[i + 2] & 0x3Fu); i += 3;
        } else if ((c & 0xF8) == 0xF0 && cont(i + 1) && cont(i + 2) && cont(i + 3)) {
            cp = ((c & 0x07u) << 18) | ((p[i + 1] & 0x3Fu) << 12) | ((p[i + 2] & 0x3Fu) << 6) | (p[i + 3] & 0x3Fu); i += 4;
        } else { cp = c; i += 1; }
        if (cp >= 0x10000) {
            cp -= 0x10000;
            out.push_back(static_cast<uint16_t>(0xD800 + (cp >> 10)));
            out.push_back(static_cast<uint16_t>(0xDC00 + (cp & 0x3FF)));
        } else {
            out.push_back(static_cast<uint16_t>(cp));
        }
    }
    return out;
}

// java.lang.String.compareTo (LagBasedPartitionAssignor.java:259).
inline int java_compare_to(const std::vector<uint16_t>& a, const std::vector<uint16_t>& b) {
    const size_t n = a.size() < b.size() ? a.size() : b.size();
    for (size_t i = 0; i < n; ++i)
        if (a[i] != b[i]) return static_cast<int>(a[i]) - static_cast<int>(b[i]);
    return static_cast<int>(a.size()) - static_cast<int>(b.size());
}

// java.lang.String.hashCode.
inline uint32_t java_string_hash(const std::string& s) {
    uint32_t h = 0;
    for (uint16_t u : utf16_units(s)) h = 31u * h + u;
    return h;
}

}  // namespace lagassign_host

#include "java_hashmap_order.hpp"   // JavaHashMapOrder: iteration order of the HashMap built by consumersPerTopic, tree bins included
