// tzdb.h -- flattened UTC-offset table of one IANA time zone (see tzdb.cc)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace auron {

struct TzTable {
    std::vector<int64_t> trans;   // utc seconds of the transitions, ascending
    std::vector<int32_t> offs;    // offs[0]: offset (seconds east) before trans[0]; offs[i + 1]: from trans[i] on
    int32_t offset_at(int64_t utc_second) const {
        size_t lo = 0, hi = trans.size();   // first transition > utc_second
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (trans[mid] <= utc_second) lo = mid + 1;
            else hi = mid;
        }
        return offs[lo];
    }
};
// false when `name` is not a zone of the tz database (chrono-tz's `parse::<Tz>()` failing, spark_dates.rs:97-102)
bool load_tz_table(const std::string& name, TzTable* out);

}  // namespace auron
