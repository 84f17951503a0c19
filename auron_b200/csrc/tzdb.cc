// tzdb.cc -- UTC-offset tables of IANA time zones for the device (E4: datafusion-ext-functions/src/spark_dates.rs:93-110,
// 200-227 resolves `s.parse::<chrono_tz::Tz>()` and asks the zone for its offset at every instant).
//
// The zone is read from the system tz database (TZif, RFC 8536: /usr/share/zoneinfo/<name>) on the host and flattened into
// { utc transition second[i], offset after it }, i ascending; instants past the last stored transition follow the POSIX TZ
// rule of the file's footer, expanded year by year up to 2200.  The expression VM binary-searches the table per row.
#include "tzdb.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>

namespace auron {
namespace {

int64_t be(const uint8_t* p, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v = (v << 8) | p[i];
    if (n == 4) return (int64_t)(int32_t)(uint32_t)v;
    return (int64_t)v;
}

int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (int64_t)doe - 719468;
}
int year_of_second(int64_t s) {
    int64_t z = (s >= 0 ? s : s - 86399) / 86400 + 719468;
    const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = (unsigned)(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const int64_t y = (int64_t)yoe + era * 400;
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doy + 2) / 153;
    return (int)(y + (mp >= 10));
}

// ---- POSIX TZ string ("EST5EDT,M3.2.0,M11.1.0"; RFC 8536 section 3.3.1)
struct Rule {
    int kind = 0;   // 0 = Mm.w.d, 1 = Jn (1..365, no leap day), 2 = n (0..365)
    int m = 0, w = 0, d = 0, n = 0;
    int32_t time = 7200;
};
struct Posix {
    bool has_dst = false;
    int32_t std_off = 0, dst_off = 0;   // seconds EAST of UTC
    Rule start, end;
};
bool skip_name(const char*& p) {
    if (*p == '<') {
        while (*p && *p != '>') p++;
        if (*p != '>') return false;
        p++;
        return true;
    }
    const char* q = p;
    while ((*p >= 'A' && *p <= 'Z') || (*p >= 'a' && *p <= 'z')) p++;
    return p - q >= 3;
}
bool parse_hms(const char*& p, int32_t* out) {   // [+-]hh[:mm[:ss]]
    int sign = 1;
    if (*p == '+') p++;
    else if (*p == '-') {
        sign = -1;
        p++;
    }
    if (*p < '0' || *p > '9') return false;
    int32_t v[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) {
        int x = 0, digits = 0;
        while (*p >= '0' && *p <= '9' && digits < 3) {
            x = x * 10 + (*p - '0');
            p++;
            digits++;
        }
        if (!digits) return false;
        v[k] = x;
        if (*p != ':') break;
        p++;
    }
    *out = sign * (v[0] * 3600 + v[1] * 60 + v[2]);
    return true;
}
bool parse_rule(const char*& p, Rule* r) {
    if (*p == 'M') {
        p++;
        r->kind = 0;
        if (sscanf(p, "%d.%d.%d", &r->m, &r->w, &r->d) != 3) return false;
        while (*p && *p != '/' && *p != ',') p++;
    } else {
        r->kind = *p == 'J' ? 1 : 2;
        if (*p == 'J') p++;
        r->n = (int)strtol(p, const_cast<char**>(&p), 10);
    }
    r->time = 7200;
    if (*p == '/') {
        p++;
        if (!parse_hms(p, &r->time)) return false;
    }
    return true;
}
bool parse_posix(const std::string& s, Posix* out) {
    const char* p = s.c_str();
    if (!skip_name(p)) return false;
    int32_t west;
    if (!parse_hms(p, &west)) return false;
    out->std_off = -west;
    if (!*p) return true;   // no daylight saving
    if (!skip_name(p)) return false;
    out->dst_off = out->std_off + 3600;
    if (*p && *p != ',') {
        if (!parse_hms(p, &west)) return false;
        out->dst_off = -west;
    }
    if (*p != ',') return false;
    p++;
    if (!parse_rule(p, &out->start) || *p != ',') return false;
    p++;
    if (!parse_rule(p, &out->end)) return false;
    out->has_dst = true;
    return true;
}
// local wall-clock second (as if UTC) at which the rule fires in `year`
int64_t rule_local_second(const Rule& r, int year) {
    int64_t day;
    if (r.kind == 0) {
        const int64_t first = days_from_civil(year, (unsigned)r.m, 1);
        int wd = (int)(((first % 7) + 11) % 7);   // 1970-01-01 is a Thursday (4); 0 = Sunday
        int delta = (r.d - wd + 7) % 7;
        day = first + delta + 7 * (r.w - 1);
        const unsigned ml[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
        unsigned len = ml[r.m - 1] + ((r.m == 2 && (year % 4 == 0 && (year % 100 != 0 || year % 400 == 0))) ? 1u : 0u);
        while (day >= first + (int64_t)len) day -= 7;   // w == 5: the last such weekday
    } else if (r.kind == 1) {
        const bool leap = year % 4 == 0 && (year % 100 != 0 || year % 400 == 0);
        day = days_from_civil(year, 1, 1) + (r.n - 1) + ((leap && r.n >= 60) ? 1 : 0);
    } else {
        day = days_from_civil(year, 1, 1) + r.n;
    }
    return day * 86400 + r.time;
}

}  // namespace

bool load_tz_table(const std::string& name, TzTable* out) {
    // chrono-tz knows the IANA names only: no paths, no "posix/" / "right/" variants, no special files
    if (name.empty() || name[0] == '/' || name.find("..") != std::string::npos || name.rfind("posix/", 0) == 0 || name.rfind("right/", 0) == 0 ||
        name == "posixrules" || name == "localtime" || name == "Factory")
        return false;
    for (char c : name)
        if (!((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '/' || c == '_' || c == '-' || c == '+')) return false;
    const char* dir = getenv("AURON_ZONEINFO");
    std::ifstream f(std::string(dir ? dir : "/usr/share/zoneinfo") + "/" + name, std::ios::binary);
    if (!f) return false;
    std::vector<uint8_t> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (b.size() < 44 || memcmp(b.data(), "TZif", 4) != 0) return false;
    size_t pos = 0;
    int width = 4;
    auto header = [&](size_t at, int64_t c[6]) {
        for (int i = 0; i < 6; i++) c[i] = be(b.data() + at + 20 + 4 * i, 4);   // isutcnt isstdcnt leapcnt timecnt typecnt charcnt
    };
    int64_t c[6];
    header(0, c);
    if (b[4] >= '2') {   // skip the 32-bit block, use the 64-bit one
        pos = 44 + (size_t)(c[3] * 4 + c[3] + c[4] * 6 + c[5] + c[2] * 8 + c[1] + c[0]);
        if (pos + 44 > b.size() || memcmp(b.data() + pos, "TZif", 4) != 0) return false;
        header(pos, c);
        width = 8;
    }
    const int64_t timecnt = c[3], typecnt = c[4];
    size_t p = pos + 44;
    const size_t need = (size_t)(timecnt * width + timecnt + typecnt * 6 + c[5] + c[2] * (width + 4) + c[1] + c[0]);
    if (typecnt < 1 || p + need > b.size()) return false;
    const uint8_t* times = b.data() + p;
    const uint8_t* idx = times + timecnt * width;
    const uint8_t* types = idx + timecnt;
    auto utoff = [&](int t) { return (int32_t)be(types + 6 * t, 4); };
    out->trans.clear();
    out->offs.clear();
    out->offs.push_back(utoff(0));   // before the first transition: time type 0 (RFC 8536 section 3.2)
    for (int64_t i = 0; i < timecnt; i++) {
        if (idx[i] >= typecnt) return false;
        const int64_t t = be(times + i * width, width);
        if (!out->trans.empty() && t <= out->trans.back()) continue;
        out->trans.push_back(t);
        out->offs.push_back(utoff(idx[i]));
    }
    // footer: "\n" POSIX-TZ "\n" (version 2+)
    if (width == 8) {
        size_t fpos = p + need;
        if (fpos < b.size() && b[fpos] == '\n') {
            size_t e = fpos + 1;
            while (e < b.size() && b[e] != '\n') e++;
            Posix px;
            const std::string tzs((const char*)b.data() + fpos + 1, e - fpos - 1);
            if (!tzs.empty() && parse_posix(tzs, &px)) {
                if (!px.has_dst) {
                    if (out->trans.empty()) out->offs[0] = px.std_off;
                } else {
                    const int y0 = out->trans.empty() ? 1900 : year_of_second(out->trans.back());
                    std::vector<std::pair<int64_t, int32_t>> ext;
                    for (int y = y0; y <= 2200; y++) {
                        ext.emplace_back(rule_local_second(px.start, y) - px.std_off, px.dst_off);   // DST starts: wall clock in standard time
                        ext.emplace_back(rule_local_second(px.end, y) - px.dst_off, px.std_off);     // DST ends: wall clock in daylight time
                    }
                    std::sort(ext.begin(), ext.end());
                    for (auto& t : ext) {
                        if (!out->trans.empty() && t.first <= out->trans.back()) continue;
                        if (out->offs.back() == t.second) continue;   // already in that state
                        out->trans.push_back(t.first);
                        out->offs.push_back(t.second);
                    }
                }
            }
        }
    }
    return true;
}

}  // namespace auron
