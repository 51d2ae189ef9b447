// parse_vw.cpp -- see parse_vw.h. The whole file is read with one read() call and scanned in place: the model
// directory holds 3 x 80 882 of these at iDASH scale, so per-line stdio calls are what the loader's time goes to.
#include "parse_vw.h"

#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <fcntl.h>
#include <unistd.h>

static inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

std::vector<std::pair<std::string, int32_t>> read_lines(const std::string &file_name) {
    const int fd = open(file_name.c_str(), O_RDONLY);
    if (fd < 0) {
        fprintf(stderr, "Cannot open file '%s'\n", file_name.c_str());
        abort();
    }
    std::string buf;
    char chunk[8192];
    for (;;) {
        const ssize_t n = ::read(fd, chunk, sizeof(chunk));
        if (n < 0 && errno == EINTR) continue;
        if (n <= 0) break;
        buf.append(chunk, (size_t) n);
    }
    close(fd);

    std::vector<std::pair<std::string, int32_t>> out;
    const char *p = buf.data(), *end = p + buf.size();
    while (p < end) {
        const char *eol = static_cast<const char *>(memchr(p, '\n', (size_t) (end - p)));
        const char *line_end = eol ? eol : end;
        // "%s": skip white space, take the run of non-space characters
        const char *q = p;
        while (q < line_end && is_space(*q)) ++q;
        const char *name = q;
        while (q < line_end && !is_space(*q)) ++q;
        if (q > name) {
            std::string key(name, (size_t) (q - name));
            // " %d": skip white space, optional sign, decimal digits; stops at the '.' of "-107.0"
            while (q < line_end && is_space(*q)) ++q;
            bool neg = false;
            if (q < line_end && (*q == '+' || *q == '-')) { neg = *q == '-'; ++q; }
            if (q >= line_end || *q < '0' || *q > '9') {
                // the reference would store an uninitialised int here; refuse instead of inventing a coefficient
                fprintf(stderr, "Cannot parse a coefficient for '%s' in file '%s'\n", key.c_str(), file_name.c_str());
                abort();
            }
            long long v = 0;
            while (q < line_end && *q >= '0' && *q <= '9') { if (v < (1LL << 40)) v = v * 10 + (*q - '0'); ++q; }
            if (neg) v = -v;
            if (v > INT_MAX) v = INT_MAX;           // glibc's %d saturates
            if (v < INT_MIN) v = INT_MIN;
            out.emplace_back(std::move(key), (int32_t) v);
        }
        p = eol ? eol + 1 : end;
    }
    return out;
}

std::unordered_map<std::string, int32_t> read(const std::string &file_name) {
    std::unordered_map<std::string, int32_t> coefs;
    for (auto &kv : read_lines(file_name)) coefs[kv.first] = kv.second;
    return coefs;
}
