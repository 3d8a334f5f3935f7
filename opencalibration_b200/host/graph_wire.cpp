// graph.json wire format of the matching path -- see graph_wire.hpp for the reference citations.
#include "graph_wire.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstring>
#include <random>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>

using namespace opencalibration;

namespace ocb_host
{
namespace wire
{
// ---------------------------------------------------------------------------------------------------------------
// base64 (src/io/base64.c)
// ---------------------------------------------------------------------------------------------------------------
namespace
{
const char kAlphabet[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

struct SextetTable
{
    uint8_t v[256];
    SextetTable()
    {
        std::memset(v, 64, sizeof(v)); // 64 = not in the alphabet (base64.c:82 pr2six)
        for (int i = 0; i < 64; i++)
            v[uint8_t(kAlphabet[i])] = uint8_t(i);
    }
};
const SextetTable kSextet;

// base64.c:180-208 (Base64encode without the trailing NUL)
size_t encode_into(const uint8_t *b, size_t n, char *out)
{
    char *p = out;
    size_t i = 0;
    for (; i + 2 < n; i += 3)
    {
        *p++ = kAlphabet[b[i] >> 2];
        *p++ = kAlphabet[((b[i] & 0x3) << 4) | (b[i + 1] >> 4)];
        *p++ = kAlphabet[((b[i + 1] & 0xF) << 2) | (b[i + 2] >> 6)];
        *p++ = kAlphabet[b[i + 2] & 0x3F];
    }
    if (i < n)
    {
        *p++ = kAlphabet[b[i] >> 2];
        if (i + 1 == n)
        {
            *p++ = kAlphabet[(b[i] & 0x3) << 4];
            *p++ = '=';
        }
        else
        {
            *p++ = kAlphabet[((b[i] & 0x3) << 4) | (b[i + 1] >> 4)];
            *p++ = kAlphabet[(b[i + 1] & 0xF) << 2];
        }
        *p++ = '=';
    }
    return size_t(p - out);
}

// base64.c:127-170: the valid prefix decides the length; 4 sextets -> 3 bytes, 3 -> 2, 2 -> 1, 1 -> dropped
size_t decode_into(const char *text, size_t n, uint8_t *out, size_t cap)
{
    size_t valid = 0;
    while (valid < n && kSextet.v[uint8_t(text[valid])] < 64)
        valid++;
    static const size_t tail_bytes[4] = {0, 0, 1, 2};
    const size_t n_out = valid / 4 * 3 + tail_bytes[valid % 4];
    if (n_out > cap)
        return size_t(-1);
    const uint8_t *s = reinterpret_cast<const uint8_t *>(text);
    size_t o = 0;
    for (size_t i = 0; i + 1 < valid && o < n_out; i += 4)
    {
        const uint32_t a = kSextet.v[s[i]], b = kSextet.v[s[i + 1]];
        const uint32_t c = i + 2 < valid ? kSextet.v[s[i + 2]] : 0, d = i + 3 < valid ? kSextet.v[s[i + 3]] : 0;
        out[o++] = uint8_t(a << 2 | b >> 4);
        if (o < n_out && i + 2 < valid)
            out[o++] = uint8_t(b << 4 | c >> 2);
        if (o < n_out && i + 3 < valid)
            out[o++] = uint8_t(c << 6 | d);
    }
    return n_out;
}
} // namespace

std::string base64_encode(const void *bytes, size_t n)
{
    std::string out((n + 2) / 3 * 4, '\0');
    out.resize(encode_into(static_cast<const uint8_t *>(bytes), n, out.data()));
    return out;
}
std::string base64_decode(const char *text, size_t n)
{
    std::string out((n + 3) / 4 * 3, '\0');
    out.resize(decode_into(text, n, reinterpret_cast<uint8_t *>(out.data()), out.size()));
    return out;
}

// bitset_to_bytes (serialize_MeasurementGraph.cpp:20-27): bit j -> byte j>>3, bit j&7. The bitset's memory image
// is little-endian u64 words, so the wire bytes are simply the first 61 bytes of the row (little-endian host).
void descriptor_row_to_base64(const uint64_t row[8], char out[DESCRIPTOR_BASE64_CHARS])
{
    uint8_t bytes[64];
    std::memcpy(bytes, row, 64);
    bytes[60] &= 0x3F; // bits 486..487 are not part of the descriptor
    encode_into(bytes, DESCRIPTOR_WIRE_BYTES, out);
}
// bitset_from_bytes (deserialize_MeasurementGraph.cpp:17-24)
bool descriptor_row_from_base64(const char *text, size_t n, uint64_t row[8])
{
    uint8_t bytes[64] = {0};
    if (decode_into(text, n, bytes, 63) != DESCRIPTOR_WIRE_BYTES)
        return false;
    bytes[60] &= 0x3F; // only 486 bits are read back (:21)
    bytes[61] = bytes[62] = bytes[63] = 0;
    std::memcpy(row, bytes, 64);
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// number formatting: Grisu2 (Loitsch 2010) + the layout rules of rapidjson's Prettify (internal/dtoa.h)
// ---------------------------------------------------------------------------------------------------------------
namespace
{
struct Fp
{
    uint64_t f;
    int e;
};
const Fp kPow10[87] = {
#include "pow10_table.inc"
};
const uint64_t kHidden = uint64_t(1) << 52;

inline Fp fp_mul(Fp a, Fp b)
{
    const unsigned __int128 p = (unsigned __int128)a.f * b.f;
    uint64_t h = uint64_t(p >> 64);
    if (uint64_t(p) >> 63) // round half up on the discarded half
        h++;
    return {h, a.e + b.e + 64};
}
inline Fp fp_normalize(Fp a)
{
    const int s = __builtin_clzll(a.f);
    return {a.f << s, a.e - s};
}

// weeds the last digit towards the exact value while staying inside the rounding interval (Grisu "round_weed")
inline void round_weed(char *digits, int len, uint64_t delta, uint64_t rest, uint64_t ten_kappa, uint64_t wp_w)
{
    while (rest < wp_w && delta - rest >= ten_kappa &&
           (rest + ten_kappa < wp_w || wp_w - rest > rest + ten_kappa - wp_w))
    {
        digits[len - 1]--;
        rest += ten_kappa;
    }
}

// digits of a positive finite double: value = digits x 10^K
void grisu2(double value, char *digits, int &len, int &K)
{
    uint64_t bits;
    std::memcpy(&bits, &value, 8);
    const int biased = int((bits >> 52) & 0x7FF);
    const uint64_t sig = bits & (kHidden - 1);
    const Fp v = biased ? Fp{sig + kHidden, biased - 1075} : Fp{sig, -1074};

    // neighbours' midpoints, on a common exponent
    const Fp plus = fp_normalize({(v.f << 1) + 1, v.e - 1});
    Fp minus = (v.f == kHidden) ? Fp{(v.f << 2) - 1, v.e - 2} : Fp{(v.f << 1) - 1, v.e - 1};
    minus.f <<= minus.e - plus.e;
    minus.e = plus.e;

    // cached power bringing the binary exponent into the digit-generation window
    const double dk = (-61 - plus.e) * 0.30102999566398114 + 347;
    int k = int(dk);
    if (dk - k > 0.0)
        k++;
    const unsigned index = unsigned((k >> 3) + 1);
    K = -(-348 + int(index << 3));
    const Fp c = kPow10[index];

    const Fp W = fp_mul(fp_normalize(v), c);
    Fp Wp = fp_mul(plus, c);
    Fp Wm = fp_mul(minus, c);
    Wm.f++;
    Wp.f--;
    uint64_t delta = Wp.f - Wm.f;

    static const uint64_t p10[20] = {1ULL,
                                     10ULL,
                                     100ULL,
                                     1000ULL,
                                     10000ULL,
                                     100000ULL,
                                     1000000ULL,
                                     10000000ULL,
                                     100000000ULL,
                                     1000000000ULL,
                                     10000000000ULL,
                                     100000000000ULL,
                                     1000000000000ULL,
                                     10000000000000ULL,
                                     100000000000000ULL,
                                     1000000000000000ULL,
                                     10000000000000000ULL,
                                     100000000000000000ULL,
                                     1000000000000000000ULL,
                                     10000000000000000000ULL};
    const int shift = -Wp.e;
    const uint64_t one = uint64_t(1) << shift;
    const uint64_t wp_w = Wp.f - W.f;
    uint32_t p1 = uint32_t(Wp.f >> shift);
    uint64_t p2 = Wp.f & (one - 1);
    int kappa = 1;
    while (kappa < 9 && p1 >= p10[kappa])
        kappa++;
    len = 0;
    while (kappa > 0)
    {
        const uint32_t unit = uint32_t(p10[kappa - 1]);
        const uint32_t d = p1 / unit;
        p1 -= d * unit;
        if (d || len)
            digits[len++] = char('0' + d);
        kappa--;
        const uint64_t rest = (uint64_t(p1) << shift) + p2;
        if (rest <= delta)
        {
            K += kappa;
            round_weed(digits, len, delta, rest, p10[kappa] << shift, wp_w);
            return;
        }
    }
    for (;;)
    {
        p2 *= 10;
        delta *= 10;
        const char d = char(p2 >> shift);
        if (d || len)
            digits[len++] = char('0' + d);
        p2 &= one - 1;
        kappa--;
        if (p2 < delta)
        {
            K += kappa;
            const int index10 = -kappa;
            round_weed(digits, len, delta, p2, one, wp_w * (index10 < 20 ? p10[index10] : 0));
            return;
        }
    }
}

char *write_exponent(int K, char *p)
{
    if (K < 0)
    {
        *p++ = '-';
        K = -K;
    }
    if (K >= 100)
    {
        *p++ = char('0' + K / 100);
        K %= 100;
        *p++ = char('0' + K / 10);
        *p++ = char('0' + K % 10);
    }
    else if (K >= 10)
    {
        *p++ = char('0' + K / 10);
        *p++ = char('0' + K % 10);
    }
    else
        *p++ = char('0' + K);
    return p;
}

// digits x 10^k -> text: plain up to 21 integer digits ("12340000000.0", "12.34"), "0.001234" down to 1e-6,
// otherwise d[.ddd]e[-]x. maxDecimalPlaces is 324 in the reference writer, so nothing is ever truncated.
char *layout(char *buf, int length, int k)
{
    const int kk = length + k; // 10^(kk-1) <= v < 10^kk
    if (0 <= k && kk <= 21)
    {
        for (int i = length; i < kk; i++)
            buf[i] = '0';
        buf[kk] = '.';
        buf[kk + 1] = '0';
        return buf + kk + 2;
    }
    if (0 < kk && kk <= 21)
    {
        std::memmove(buf + kk + 1, buf + kk, size_t(length - kk));
        buf[kk] = '.';
        return buf + length + 1;
    }
    if (-6 < kk && kk <= 0)
    {
        const int offset = 2 - kk;
        std::memmove(buf + offset, buf, size_t(length));
        buf[0] = '0';
        buf[1] = '.';
        for (int i = 2; i < offset; i++)
            buf[i] = '0';
        return buf + length + offset;
    }
    if (length == 1)
    {
        buf[1] = 'e';
        return write_exponent(kk - 1, buf + 2);
    }
    std::memmove(buf + 2, buf + 1, size_t(length - 1));
    buf[1] = '.';
    buf[length + 1] = 'e';
    return write_exponent(kk - 1, buf + length + 2);
}
} // namespace

size_t format_double(double value, char *buf)
{
    if (std::isnan(value))
    {
        std::memcpy(buf, "NaN", 3);
        return 3;
    }
    if (std::isinf(value))
    {
        const char *s = value < 0 ? "-Infinity" : "Infinity";
        const size_t n = std::strlen(s);
        std::memcpy(buf, s, n);
        return n;
    }
    char *p = buf;
    if (value == 0.0)
    {
        if (std::signbit(value))
            *p++ = '-';
        std::memcpy(p, "0.0", 3);
        return size_t(p + 3 - buf);
    }
    if (value < 0)
    {
        *p++ = '-';
        value = -value;
    }
    int len, K;
    grisu2(value, p, len, K);
    return size_t(layout(p, len, K) - buf);
}

bool parse_double(const char *text, size_t n, double &value)
{
    // kParseNanAndInfFlag spellings (rapidjson reader.h ParseNumber): NaN, Inf, Infinity, with optional '-'
    const bool neg = n > 0 && text[0] == '-';
    const char *t = text + (neg ? 1 : 0);
    const size_t m = n - (neg ? 1 : 0);
    if (m == 3 && std::memcmp(t, "NaN", 3) == 0) // "-NaN" is taken too (the minus is consumed first)
    {
        value = std::numeric_limits<double>::quiet_NaN();
        return true;
    }
    if ((m == 3 && std::memcmp(t, "Inf", 3) == 0) || (m == 8 && std::memcmp(t, "Infinity", 8) == 0))
    {
        value = neg ? -std::numeric_limits<double>::infinity() : std::numeric_limits<double>::infinity();
        return true;
    }
    // JSON grammar: -? (0 | [1-9][0-9]*) (\.[0-9]+)? ([eE][+-]?[0-9]+)?   (from_chars alone is laxer: "1.", "01")
    size_t i = 0;
    auto digits = [&] {
        const size_t b = i;
        while (i < m && t[i] >= '0' && t[i] <= '9')
            i++;
        return i - b;
    };
    if (i < m && t[i] == '0')
        i++;
    else if (digits() == 0)
        return false;
    if (i < m && t[i] == '.')
    {
        i++;
        if (digits() == 0)
            return false;
    }
    if (i < m && (t[i] == 'e' || t[i] == 'E'))
    {
        i++;
        if (i < m && (t[i] == '+' || t[i] == '-'))
            i++;
        if (digits() == 0)
            return false;
    }
    if (i != m)
        return false;
    const auto r = std::from_chars(text, text + n, value); // correctly rounded = kParseFullPrecisionFlag
    if (r.ec == std::errc::result_out_of_range)
        return false; // rapidjson: kParseErrorNumberTooBig
    return r.ec == std::errc() && r.ptr == text + n;
}

// ---------------------------------------------------------------------------------------------------------------
// scanner: order-free descent over the JSON text, no DOM
// ---------------------------------------------------------------------------------------------------------------
namespace
{
struct ParseError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

class Scanner
{
  public:
    Scanner(const char *text, size_t n) : _p(text), _begin(text), _end(text + n) {}

    [[noreturn]] void fail(const std::string &what) const
    {
        throw ParseError(what + " at byte " + std::to_string(size_t(_p - _begin)));
    }
    void ws()
    {
        while (_p < _end && (*_p == ' ' || *_p == '\n' || *_p == '\r' || *_p == '\t'))
            _p++;
    }
    char peek()
    {
        ws();
        if (_p >= _end)
            fail("unexpected end of text");
        return *_p;
    }
    void expect(char c)
    {
        if (peek() != c)
            fail(std::string("expected '") + c + "'");
        _p++;
    }
    bool accept(char c)
    {
        if (peek() == c)
        {
            _p++;
            return true;
        }
        return false;
    }
    bool at_end()
    {
        ws();
        return _p >= _end;
    }

    // for (obj.begin(); obj.next(key);) { value }   -- members in document order
    void begin_object() { expect('{'); _first.push_back(true); }
    bool next_member(std::string &key)
    {
        if (accept('}'))
        {
            _first.pop_back();
            return false;
        }
        if (!_first.back())
            expect(',');
        _first.back() = false;
        string(key);
        expect(':');
        return true;
    }
    void begin_array() { expect('['); _first.push_back(true); }
    bool next_element()
    {
        if (accept(']'))
        {
            _first.pop_back();
            return false;
        }
        if (!_first.back())
            expect(',');
        _first.back() = false;
        return true;
    }

    void string(std::string &out)
    {
        const char *b, *e;
        if (raw_string(b, e))
        {
            out.assign(b, e);
            return;
        }
        unescape(b, e, out);
    }
    // string body without unescaping; returns true when it holds no escape (then [b, e) is the value itself)
    bool raw_string(const char *&b, const char *&e)
    {
        expect('"');
        b = _p;
        bool plain = true;
        while (_p < _end && *_p != '"')
        {
            if (*_p == '\\')
            {
                plain = false;
                _p++;
            }
            else if (uint8_t(*_p) < 0x20)
                fail("control character in string");
            _p++;
        }
        if (_p >= _end)
            fail("unterminated string");
        e = _p++;
        return plain;
    }
    void number_token(const char *&b, const char *&e)
    {
        ws();
        b = _p;
        while (_p < _end && (std::strchr("+-.0123456789eE", *_p) != nullptr || (*_p >= 'A' && *_p <= 'Z') ||
                             (*_p >= 'a' && *_p <= 'z')))
            _p++;
        e = _p;
        if (b == e)
            fail("expected a number");
    }
    double number() // GetDouble()
    {
        const char *b, *e;
        number_token(b, e);
        double v;
        if (!parse_double(b, size_t(e - b), v))
        {
            _p = b;
            fail("malformed number");
        }
        return v;
    }
    int64_t int64() // GetInt64(): integer tokens only
    {
        const char *b, *e;
        number_token(b, e);
        int64_t v = 0;
        const auto r = std::from_chars(b, e, v);
        if (r.ec != std::errc() || r.ptr != e)
        {
            // values in (INT64_MAX, UINT64_MAX] written by Uint64 read back through the unsigned path
            uint64_t u = 0;
            const auto ru = std::from_chars(b, e, u);
            if (ru.ec != std::errc() || ru.ptr != e)
            {
                _p = b;
                fail("expected an integer");
            }
            v = int64_t(u);
        }
        return v;
    }
    uint64_t uint64()
    {
        const char *b, *e;
        number_token(b, e);
        uint64_t v = 0;
        const auto r = std::from_chars(b, e, v);
        if (r.ec != std::errc() || r.ptr != e)
        {
            _p = b;
            fail("expected an unsigned integer");
        }
        return v;
    }
    void skip_value()
    {
        const char c = peek();
        std::string key;
        if (c == '{')
        {
            for (begin_object(); next_member(key);)
                skip_value();
        }
        else if (c == '[')
        {
            for (begin_array(); next_element();)
                skip_value();
        }
        else if (c == '"')
        {
            const char *b, *e;
            raw_string(b, e);
        }
        else if (c == 't' || c == 'f' || c == 'n')
        {
            const char *lit = c == 't' ? "true" : c == 'f' ? "false" : "null";
            const size_t n = std::strlen(lit);
            if (size_t(_end - _p) < n || std::memcmp(_p, lit, n) != 0)
                fail("unknown literal");
            _p += n;
        }
        else
            number();
    }

  private:
    static void append_utf8(uint32_t cp, std::string &out)
    {
        if (cp < 0x80)
            out.push_back(char(cp));
        else if (cp < 0x800)
        {
            out.push_back(char(0xC0 | (cp >> 6)));
            out.push_back(char(0x80 | (cp & 0x3F)));
        }
        else if (cp < 0x10000)
        {
            out.push_back(char(0xE0 | (cp >> 12)));
            out.push_back(char(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back(char(0x80 | (cp & 0x3F)));
        }
        else
        {
            out.push_back(char(0xF0 | (cp >> 18)));
            out.push_back(char(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back(char(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back(char(0x80 | (cp & 0x3F)));
        }
    }
    uint32_t hex4(const char *&s, const char *e) const
    {
        if (e - s < 4)
            fail("truncated \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++, s++)
        {
            const char c = *s;
            v <<= 4;
            if (c >= '0' && c <= '9')
                v |= uint32_t(c - '0');
            else if (c >= 'a' && c <= 'f')
                v |= uint32_t(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F')
                v |= uint32_t(c - 'A' + 10);
            else
                fail("bad \\u escape");
        }
        return v;
    }
    void unescape(const char *s, const char *e, std::string &out) const
    {
        out.clear();
        while (s < e)
        {
            if (*s != '\\')
            {
                out.push_back(*s++);
                continue;
            }
            if (++s >= e)
                fail("dangling escape");
            const char c = *s++;
            switch (c)
            {
            case '"': out.push_back('"'); break;
            case '\\': out.push_back('\\'); break;
            case '/': out.push_back('/'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'n': out.push_back('\n'); break;
            case 'r': out.push_back('\r'); break;
            case 't': out.push_back('\t'); break;
            case 'u': {
                uint32_t cp = hex4(s, e);
                if (cp >= 0xD800 && cp <= 0xDBFF)
                {
                    if (e - s < 2 || s[0] != '\\' || s[1] != 'u')
                        fail("lone surrogate");
                    s += 2;
                    const uint32_t lo = hex4(s, e);
                    if (lo < 0xDC00 || lo > 0xDFFF)
                        fail("bad surrogate pair");
                    cp = (((cp - 0xD800) << 10) | (lo - 0xDC00)) + 0x10000;
                }
                append_utf8(cp, out);
                break;
            }
            default: fail("unknown escape");
            }
        }
    }

    const char *_p, *_begin, *_end;
    std::vector<bool> _first;
};

// ---------------------------------------------------------------------------------------------------------------
// emitter: rapidjson PrettyWriter with kFormatSingleLineArray (prettywriter.h PrettyPrefix / EndObject / EndArray)
// ---------------------------------------------------------------------------------------------------------------
class Emitter
{
  public:
    explicit Emitter(std::string &out) : _out(out) {}

    void begin_object()
    {
        prefix();
        _levels.push_back({false, 0});
        _out.push_back('{');
    }
    void end_object()
    {
        const bool empty = _levels.back().count == 0;
        _levels.pop_back();
        if (!empty)
        {
            _out.push_back('\n');
            indent();
        }
        _out.push_back('}');
    }
    void begin_array()
    {
        prefix();
        _levels.push_back({true, 0});
        _out.push_back('[');
    }
    void end_array()
    {
        _levels.pop_back(); // single-line arrays: no newline before ']'
        _out.push_back(']');
    }
    void key(const char *k) { string(k, std::strlen(k)); }
    void key(const std::string &k) { string(k.data(), k.size()); }
    void string(const std::string &s) { string(s.data(), s.size()); }
    void string(const char *s, size_t n)
    {
        prefix();
        static const char hex[] = "0123456789ABCDEF";
        _out.push_back('"');
        for (size_t i = 0; i < n; i++)
        {
            const uint8_t c = uint8_t(s[i]);
            if (c == '"' || c == '\\')
            {
                _out.push_back('\\');
                _out.push_back(char(c));
            }
            else if (c < 0x20)
            {
                _out.push_back('\\');
                switch (c)
                {
                case '\b': _out.push_back('b'); break;
                case '\f': _out.push_back('f'); break;
                case '\n': _out.push_back('n'); break;
                case '\r': _out.push_back('r'); break;
                case '\t': _out.push_back('t'); break;
                default:
                    _out.append("u00");
                    _out.push_back(hex[c >> 4]);
                    _out.push_back(hex[c & 0xF]);
                }
            }
            else
                _out.push_back(char(c));
        }
        _out.push_back('"');
    }
    void int64(int64_t v)
    {
        prefix();
        char buf[24];
        const auto r = std::to_chars(buf, buf + sizeof(buf), v);
        _out.append(buf, r.ptr);
    }
    void uint64(uint64_t v)
    {
        prefix();
        char buf[24];
        const auto r = std::to_chars(buf, buf + sizeof(buf), v);
        _out.append(buf, r.ptr);
    }
    void number(double v)
    {
        prefix();
        char buf[40];
        _out.append(buf, format_double(v, buf));
    }

  private:
    struct Level
    {
        bool in_array;
        size_t count;
    };
    void indent() { _out.append(_levels.size() * 4, ' '); }
    void prefix()
    {
        if (_levels.empty())
            return;
        Level &l = _levels.back();
        if (l.in_array)
        {
            if (l.count > 0)
                _out.append(", ");
        }
        else
        {
            if (l.count == 0)
                _out.push_back('\n');
            else if (l.count % 2 == 0)
                _out.append(",\n");
            else
                _out.append(": ");
            if (l.count % 2 == 0)
                indent();
        }
        l.count++;
    }
    std::string &_out;
    std::vector<Level> _levels;
};

size_t parse_id(const std::string &s) // std::strtoull(name, &end, 10) (deserialize :57)
{
    return size_t(std::strtoull(s.c_str(), nullptr, 10));
}

template <size_t N> void read_doubles(Scanner &sc, double (&v)[N], const char *what)
{
    size_t i = 0;
    for (sc.begin_array(); sc.next_element(); i++)
    {
        const double d = sc.number();
        if (i < N)
            v[i] = d;
    }
    if (i < N)
        sc.fail(std::string("too few elements in ") + what);
}

void read_features(Scanner &sc, std::vector<feature_2d> &features)
{
    features.clear();
    std::string key;
    for (sc.begin_array(); sc.next_element();)
    {
        feature_2d f;
        unsigned seen = 0;
        for (sc.begin_object(); sc.next_member(key);)
        {
            if (key == "location")
            {
                double xy[2];
                read_doubles(sc, xy, "location");
                f.location = Eigen::Vector2d(xy[0], xy[1]);
                seen |= 1;
            }
            else if (key == "strength")
            {
                f.strength = float(sc.number());
                seen |= 2;
            }
            else if (key == "descriptor")
            {
                const char *b, *e;
                std::string unescaped;
                if (!sc.raw_string(b, e))
                    sc.fail("escape sequence in a descriptor");
                static_assert(sizeof(f.descriptor) == 64, "descriptor row");
                uint64_t row[8];
                if (!descriptor_row_from_base64(b, size_t(e - b), row))
                    sc.fail("descriptor is not 61 base64-coded bytes");
                std::memcpy(static_cast<void *>(&f.descriptor), row, 64);
                seen |= 4;
            }
            else
                sc.skip_value();
        }
        if (seen != 7)
            sc.fail("feature without location/strength/descriptor");
        features.push_back(f);
    }
}

void read_node(Scanner &sc, GraphNode &node)
{
    std::string key, k2, k3;
    unsigned seen = 0;
    bool has_sparse = false;
    for (sc.begin_object(); sc.next_member(key);)
    {
        if (key == "path")
            sc.string(node.path), seen |= 1;
        else if (key == "position")
            read_doubles(sc, node.position, "position"), seen |= 2;
        else if (key == "orientation")
            read_doubles(sc, node.orientation_xyzw, "orientation"), seen |= 4;
        else if (key == "thumbnail")
            sc.string(node.thumbnail_base64), seen |= 8;
        else if (key == "model")
        {
            seen |= 16;
            for (sc.begin_object(); sc.next_member(k2);)
            {
                if (k2 == "id")
                    node.model_id = sc.int64();
                else if (k2 == "dimensions")
                {
                    size_t i = 0;
                    for (sc.begin_array(); sc.next_element(); i++)
                    {
                        const int64_t v = sc.int64();
                        if (i == 0)
                            node.model.pixels_cols = size_t(v);
                        else if (i == 1)
                            node.model.pixels_rows = size_t(v);
                    }
                }
                else if (k2 == "focal_length")
                    node.model.focal_length_pixels = sc.number();
                else if (k2 == "principal")
                {
                    double v[2];
                    read_doubles(sc, v, "principal");
                    node.model.principle_point = Eigen::Vector2d(v[0], v[1]);
                }
                else if (k2 == "radial_distortion")
                {
                    double v[3];
                    read_doubles(sc, v, "radial_distortion");
                    node.model.radial_distortion = Eigen::Vector3d(v[0], v[1], v[2]);
                }
                else if (k2 == "tangential_distortion")
                {
                    double v[2];
                    read_doubles(sc, v, "tangential_distortion");
                    node.model.tangential_distortion = Eigen::Vector2d(v[0], v[1]);
                }
                else if (k2 == "projection")
                {
                    sc.string(node.projection);
                    node.model.projection_type =
                        node.projection == "planar" ? ProjectionType::PLANAR : ProjectionType::UNKNOWN;
                }
                else
                    sc.skip_value();
            }
        }
        else if (key == "edges")
        {
            seen |= 32;
            node.edges.clear();
            for (sc.begin_array(); sc.next_element();)
            {
                sc.string(k2);
                node.edges.push_back(parse_id(k2));
            }
        }
        else if (key == "metadata")
        {
            seen |= 64;
            for (sc.begin_object(); sc.next_member(k2);)
            {
                if (k2 == "camera_info")
                {
                    CameraInfo &c = node.camera_info;
                    for (sc.begin_object(); sc.next_member(k3);)
                    {
                        if (k3 == "dimensions")
                        {
                            size_t i = 0;
                            for (sc.begin_array(); sc.next_element(); i++)
                            {
                                const int64_t v = sc.int64();
                                (i == 0 ? c.width_px : c.height_px) = uint64_t(v);
                            }
                        }
                        else if (k3 == "focal_length_px")
                            c.focal_length_px = sc.number();
                        else if (k3 == "principal")
                            read_doubles(sc, c.principal_point_px, "principal");
                        else if (k3 == "make")
                            sc.string(c.make);
                        else if (k3 == "model")
                            sc.string(c.model);
                        else if (k3 == "serial_no")
                            sc.string(c.serial_no);
                        else if (k3 == "lens_make")
                            sc.string(c.lens_make);
                        else if (k3 == "lens_model")
                            sc.string(c.lens_model);
                        else
                            sc.skip_value();
                    }
                }
                else if (k2 == "capture_info")
                {
                    CaptureInfo &c = node.capture_info;
                    for (sc.begin_object(); sc.next_member(k3);)
                    {
                        if (k3 == "latitude")
                            c.latitude = sc.number();
                        else if (k3 == "longitude")
                            c.longitude = sc.number();
                        else if (k3 == "altitude")
                            c.altitude = sc.number();
                        else if (k3 == "relative_altitude")
                            c.relative_altitude = sc.number();
                        else if (k3 == "roll")
                            c.roll = sc.number();
                        else if (k3 == "pitch")
                            c.pitch = sc.number();
                        else if (k3 == "yaw")
                            c.yaw = sc.number();
                        else if (k3 == "accuracy_xy")
                            c.accuracy_xy = sc.number();
                        else if (k3 == "accuracy_z")
                            c.accuracy_z = sc.number();
                        else if (k3 == "datum")
                            sc.string(c.datum);
                        else if (k3 == "timestamp")
                            sc.string(c.timestamp);
                        else if (k3 == "datestamp")
                            sc.string(c.datestamp);
                        else
                            sc.skip_value();
                    }
                }
                else
                    sc.skip_value();
            }
        }
        else if (key == "features")
            read_features(sc, node.features), seen |= 128;
        else if (key == "num_sparse_features")
            node.num_sparse_features = size_t(sc.uint64()), has_sparse = true;
        else
            sc.skip_value();
    }
    if (seen != 255)
        sc.fail("image node " + std::to_string(node.id) + " lacks a member the reference reads");
    if (!has_sparse)
        node.num_sparse_features = node.features.size(); // deserialize :180-187
    // The reference trusts the file (link_stage.cpp:63-65 then reads features[0 .. num_sparse) unchecked); a corrupt
    // or crafted checkpoint must not become an out-of-bounds read of descriptor memory here.
    if (node.num_sparse_features > node.features.size())
        sc.fail("image node " + std::to_string(node.id) + ": num_sparse_features exceeds the number of features");
}

void read_edge(Scanner &sc, GraphEdge &edge)
{
    std::string key, k2;
    unsigned seen = 0;
    camera_relations &rel = edge.relations;
    for (sc.begin_object(); sc.next_member(key);)
    {
        if (key == "source")
            sc.string(k2), edge.source = parse_id(k2), seen |= 1;
        else if (key == "dest")
            sc.string(k2), edge.dest = parse_id(k2), seen |= 2;
        else if (key == "matches")
        {
            seen |= 4;
            rel.matches.clear();
            for (sc.begin_array(); sc.next_element();)
            {
                feature_match m{0, 0, 0.0};
                sc.begin_array();
                if (!sc.next_element())
                    sc.fail("empty match");
                m.feature_index_1 = size_t(sc.int64());
                if (!sc.next_element())
                    sc.fail("short match");
                m.feature_index_2 = size_t(sc.int64());
                if (!sc.next_element())
                    sc.fail("short match");
                m.distance = sc.number();
                while (sc.next_element())
                    sc.skip_value();
                rel.matches.push_back(m);
            }
        }
        else if (key == "inlier_matches")
        {
            seen |= 8;
            rel.inlier_matches.clear();
            for (sc.begin_array(); sc.next_element();)
            {
                feature_match_denormalized m;
                double px[2];
                size_t i = 0;
                for (sc.begin_array(); sc.next_element(); i++)
                {
                    switch (i)
                    {
                    case 0:
                        read_doubles(sc, px, "pixel_1");
                        m.pixel_1 = Eigen::Vector2d(px[0], px[1]);
                        break;
                    case 1:
                        read_doubles(sc, px, "pixel_2");
                        m.pixel_2 = Eigen::Vector2d(px[0], px[1]);
                        break;
                    case 2: m.feature_index_1 = size_t(sc.int64()); break;
                    case 3: m.feature_index_2 = size_t(sc.int64()); break;
                    case 4: m.match_index = size_t(sc.int64()); break;
                    default: sc.skip_value();
                    }
                }
                if (i < 5)
                    sc.fail("short inlier match");
                rel.inlier_matches.push_back(m);
            }
        }
        else if (key == "relation")
        {
            seen |= 16;
            double r[9];
            read_doubles(sc, r, "relation");
            for (int i = 0; i < 3; i++) // row-major on the wire (:530-537)
                for (int j = 0; j < 3; j++)
                    rel.ransac_relation(i, j) = r[i * 3 + j];
        }
        else if (key == "relation_type")
        {
            seen |= 32;
            sc.string(k2);
            rel.relationType = k2 == "homography"           ? camera_relations::RelationType::HOMOGRAPHY
                               : k2 == "fundamental_matrix" ? camera_relations::RelationType::FUNDAMENTAL_MATRIX
                                                            : camera_relations::RelationType::UNKNOWN;
        }
        else if (key == "relative_pose")
        {
            seen |= 64;
            size_t i = 0;
            for (sc.begin_array(); sc.next_element(); i++)
            {
                if (i >= rel.relative_poses.size())
                    sc.fail("more than 4 relative poses");
                decomposed_pose &pose = rel.relative_poses[i];
                for (sc.begin_object(); sc.next_member(k2);)
                {
                    if (k2 == "score")
                        pose.score = int(sc.int64());
                    else if (k2 == "orientation")
                    {
                        double q[4];
                        read_doubles(sc, q, "orientation");
                        for (int j = 0; j < 4; j++)
                            pose.orientation.coeffs()(j) = q[j];
                    }
                    else if (k2 == "position")
                    {
                        double t[3];
                        read_doubles(sc, t, "position");
                        for (int j = 0; j < 3; j++)
                            pose.position(j) = t[j];
                    }
                    else
                        sc.skip_value();
                }
            }
            edge.n_relative_poses = i;
        }
        else
            sc.skip_value();
    }
    if (seen != 127)
        sc.fail("edge " + std::to_string(edge.id) + " lacks a member the reference reads");
}

void write_doubles(Emitter &w, const double *v, size_t n)
{
    w.begin_array();
    for (size_t i = 0; i < n; i++)
        w.number(v[i]);
    w.end_array();
}

void write_node(Emitter &w, const GraphNode &node)
{
    w.key(std::to_string(node.id));
    w.begin_object();
    w.key("path");
    w.string(node.path);
    w.key("position");
    write_doubles(w, node.position, 3);
    w.key("orientation");
    write_doubles(w, node.orientation_xyzw, 4);
    w.key("thumbnail");
    w.string(node.thumbnail_base64);
    w.key("model");
    w.begin_object();
    {
        w.key("id");
        w.int64(node.model_id);
        w.key("dimensions");
        w.begin_array();
        w.uint64(node.model.pixels_cols);
        w.uint64(node.model.pixels_rows);
        w.end_array();
        w.key("focal_length");
        w.number(node.model.focal_length_pixels);
        w.key("principal");
        write_doubles(w, node.model.principle_point.data(), 2);
        w.key("radial_distortion");
        write_doubles(w, node.model.radial_distortion.data(), 3);
        w.key("tangential_distortion");
        write_doubles(w, node.model.tangential_distortion.data(), 2);
        w.key("projection");
        w.string(node.model.projection_type == ProjectionType::PLANAR ? "planar" : "UNKNOWN");
    }
    w.end_object();
    w.key("edges");
    w.begin_array();
    {
        std::vector<size_t> sorted(node.edges);
        std::sort(sorted.begin(), sorted.end());
        sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end()); // a set in the reference
        for (size_t id : sorted)
            w.string(std::to_string(id));
    }
    w.end_array();
    w.key("metadata");
    w.begin_object();
    {
        w.key("camera_info");
        w.begin_object();
        const CameraInfo &c = node.camera_info;
        w.key("dimensions");
        w.begin_array();
        w.uint64(c.width_px);
        w.uint64(c.height_px);
        w.end_array();
        w.key("focal_length_px");
        w.number(c.focal_length_px);
        w.key("principal");
        write_doubles(w, c.principal_point_px, 2);
        w.key("make");
        w.string(c.make);
        w.key("model");
        w.string(c.model);
        w.key("serial_no");
        w.string(c.serial_no);
        w.key("lens_make");
        w.string(c.lens_make);
        w.key("lens_model");
        w.string(c.lens_model);
        w.end_object();

        w.key("capture_info");
        w.begin_object();
        const CaptureInfo &p = node.capture_info;
        w.key("latitude");
        w.number(p.latitude);
        w.key("longitude");
        w.number(p.longitude);
        w.key("altitude");
        w.number(p.altitude);
        w.key("relative_altitude");
        w.number(p.relative_altitude);
        w.key("roll");
        w.number(p.roll);
        w.key("pitch");
        w.number(p.pitch);
        w.key("yaw");
        w.number(p.yaw);
        w.key("accuracy_xy");
        w.number(p.accuracy_xy);
        w.key("accuracy_z");
        w.number(p.accuracy_z);
        w.key("datum");
        w.string(p.datum);
        w.key("timestamp");
        w.string(p.timestamp);
        w.key("datestamp");
        w.string(p.datestamp);
        w.end_object();
    }
    w.end_object();

    w.key("features");
    w.begin_array();
    char b64[DESCRIPTOR_BASE64_CHARS];
    for (const feature_2d &f : node.features)
    {
        w.begin_object();
        w.key("location");
        write_doubles(w, f.location.data(), 2);
        w.key("strength");
        w.number(double(f.strength));
        w.key("descriptor");
        uint64_t row[8];
        std::memcpy(row, static_cast<const void *>(&f.descriptor), 64);
        descriptor_row_to_base64(row, b64);
        w.string(b64, DESCRIPTOR_BASE64_CHARS);
        w.end_object();
    }
    w.end_array();
    w.key("num_sparse_features");
    w.uint64(node.num_sparse_features);
    w.end_object();
}

void write_edge(Emitter &w, const GraphEdge &edge)
{
    const camera_relations &rel = edge.relations;
    w.key(std::to_string(edge.id));
    w.begin_object();
    w.key("source");
    w.string(std::to_string(edge.source));
    w.key("dest");
    w.string(std::to_string(edge.dest));
    w.key("matches");
    w.begin_array();
    for (const feature_match &m : rel.matches)
    {
        w.begin_array();
        w.int64(int64_t(m.feature_index_1));
        w.int64(int64_t(m.feature_index_2));
        w.number(m.distance);
        w.end_array();
    }
    w.end_array();
    w.key("inlier_matches");
    w.begin_array();
    for (const feature_match_denormalized &m : rel.inlier_matches)
    {
        w.begin_array();
        write_doubles(w, m.pixel_1.data(), 2);
        write_doubles(w, m.pixel_2.data(), 2);
        w.int64(int64_t(m.feature_index_1));
        w.int64(int64_t(m.feature_index_2));
        w.int64(int64_t(m.match_index));
        w.end_array();
    }
    w.end_array();
    w.key("relation");
    w.begin_array();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            w.number(rel.ransac_relation(i, j));
    w.end_array();
    w.key("relation_type");
    w.string(rel.relationType == camera_relations::RelationType::HOMOGRAPHY            ? "homography"
             : rel.relationType == camera_relations::RelationType::FUNDAMENTAL_MATRIX ? "fundamental_matrix"
                                                                                       : "UNKNOWN");
    w.key("relative_pose");
    w.begin_array();
    for (const decomposed_pose &pose : rel.relative_poses)
    {
        w.begin_object();
        w.key("score");
        w.int64(pose.score);
        w.key("orientation");
        w.begin_array();
        for (int i = 0; i < 4; i++)
            w.number(pose.orientation.coeffs()(i));
        w.end_array();
        w.key("position");
        w.begin_array();
        for (int i = 0; i < 3; i++)
            w.number(pose.position(i));
        w.end_array();
        w.end_object();
    }
    w.end_array();
    w.end_object();
}
} // namespace

bool read_graph(const char *json, size_t n, GraphDocument &graph, std::string *error)
{
    GraphDocument doc;
    try
    {
        Scanner sc(json, n);
        if (sc.at_end() || sc.peek() != '{')
            throw ParseError("the document is not a JSON object");
        // The reference looks members up by name after a DOM parse (deserialize :49-56), so "version" may follow
        // "nodes": everything is read first and the version decides at the end.
        bool version_ok = false, has_nodes = false, has_edges = false;
        std::string key, id;
        for (sc.begin_object(); sc.next_member(key);)
        {
            if (key == "version")
            {
                const char c = sc.peek();
                if (c == '-' || (c >= '0' && c <= '9'))
                {
                    const char *b, *e;
                    sc.number_token(b, e);
                    int64_t v = 0;
                    const auto r = std::from_chars(b, e, v);
                    version_ok = r.ec == std::errc() && r.ptr == e && v == 1;
                }
                else
                    sc.skip_value();
            }
            else if (key == "nodes")
            {
                has_nodes = true;
                doc.nodes.clear();
                // deserialize :84-110: camera models are shared by model id -- a node whose model id was seen
                // earlier in the document gets that earlier model, whatever its own "model" members say
                std::unordered_map<int64_t, size_t> first_with_model;
                for (sc.begin_object(); sc.next_member(id);)
                {
                    doc.nodes.emplace_back();
                    GraphNode &node = doc.nodes.back();
                    node.id = parse_id(id);
                    read_node(sc, node);
                    const auto seen = first_with_model.emplace(node.model_id, doc.nodes.size() - 1);
                    if (!seen.second)
                        node.model = doc.nodes[seen.first->second].model;
                }
            }
            else if (key == "edges")
            {
                has_edges = true;
                doc.edges.clear();
                for (sc.begin_object(); sc.next_member(id);)
                {
                    doc.edges.emplace_back();
                    doc.edges.back().id = parse_id(id);
                    read_edge(sc, doc.edges.back());
                }
            }
            else
                sc.skip_value();
        }
        if (!sc.at_end())
            sc.fail("text after the document");
        if (!version_ok)
            throw ParseError("not a version-1 graph");
        if (!has_nodes || !has_edges)
            throw ParseError("graph without nodes/edges");
    }
    catch (const ParseError &e)
    {
        if (error)
            *error = e.what();
        return false;
    }
    graph = std::move(doc);
    return true;
}

void write_graph(const GraphDocument &graph, std::string &out)
{
    out.clear();
    size_t n_features = 0, n_matches = 0;
    for (const GraphNode &n : graph.nodes)
        n_features += n.features.size();
    for (const GraphEdge &e : graph.edges)
        n_matches += e.relations.matches.size() + 2 * e.relations.inlier_matches.size();
    out.reserve(4096 + graph.nodes.size() * 2048 + n_features * 260 + n_matches * 64);

    std::vector<const GraphNode *> nodes;
    for (const GraphNode &n : graph.nodes)
        nodes.push_back(&n);
    std::sort(nodes.begin(), nodes.end(), [](const GraphNode *a, const GraphNode *b) { return a->id < b->id; });
    std::vector<const GraphEdge *> edges;
    for (const GraphEdge &e : graph.edges)
        edges.push_back(&e);
    std::sort(edges.begin(), edges.end(), [](const GraphEdge *a, const GraphEdge *b) { return a->id < b->id; });

    Emitter w(out);
    w.begin_object();
    w.key("version");
    w.int64(1);
    w.key("nodes");
    w.begin_object();
    for (const GraphNode *n : nodes)
        write_node(w, *n);
    w.end_object();
    w.key("edges");
    w.begin_object();
    for (const GraphEdge *e : edges)
        write_edge(w, *e);
    w.end_object();
    w.end_object();
}

const GraphNode *GraphDocument::find_node(size_t id) const
{
    for (const GraphNode &n : nodes)
        if (n.id == id)
            return &n;
    return nullptr;
}
const GraphEdge *GraphDocument::find_edge(size_t source, size_t dest) const
{
    for (const GraphEdge &e : edges)
        if (e.source == source && e.dest == dest)
            return &e;
    return nullptr;
}

size_t GraphDocument::add_node(GraphNode node)
{
    size_t identifier = _distribution(_generator);
    while (find_node(identifier) != nullptr)
        identifier = _distribution(_generator);
    node.id = identifier;
    nodes.push_back(std::move(node));
    return identifier;
}

size_t GraphDocument::draw_edge_id(const std::unordered_set<size_t> &taken)
{
    size_t identifier = _distribution(_generator);
    while (taken.count(identifier))
        identifier = _distribution(_generator);
    return identifier;
}

size_t GraphDocument::add_edge(camera_relations relations, size_t source, size_t dest)
{
    std::unordered_set<size_t> taken;
    for (const GraphEdge &e : edges)
        taken.insert(e.id);
    GraphEdge e;
    e.id = draw_edge_id(taken);
    e.source = source;
    e.dest = dest;
    e.relations = std::move(relations);
    const size_t identifier = e.id;
    edges.push_back(std::move(e));
    for (GraphNode &n : nodes)
        if (n.id == source || n.id == dest)
            n.edges.push_back(identifier);
    return identifier;
}

LinkStats link_graph(GraphDocument &graph, const std::vector<LinkPair> &pairs_by_node_id, const LinkOptions &options)
{
    std::unordered_map<size_t, size_t> slot;
    std::vector<LinkImage> images(graph.nodes.size());
    for (size_t i = 0; i < graph.nodes.size(); i++)
    {
        const GraphNode &n = graph.nodes[i];
        slot.emplace(n.id, i);
        images[i].features = &n.features;
        images[i].num_sparse_features = n.num_sparse_features;
        images[i].model = n.model;
    }
    std::vector<LinkPair> pairs;
    pairs.reserve(pairs_by_node_id.size());
    for (const LinkPair &p : pairs_by_node_id)
    {
        const auto a = slot.find(p.image_1), b = slot.find(p.image_2);
        if (a == slot.end() || b == slot.end())
            throw std::runtime_error("link_graph: pair names a node that is not in the graph");
        pairs.push_back({a->second, b->second});
    }
    LinkStats stats;
    std::vector<camera_relations> relations = link_pairs(images, pairs, options, &stats);

    std::unordered_set<size_t> taken; // one id set for all new edges (add_edge rebuilds it per call)
    std::unordered_map<uint64_t, size_t> by_ends; // (source slot, dest slot) -> edge position
    auto ends_key = [&](size_t s, size_t d) { return uint64_t(slot.at(s)) << 32 | uint64_t(slot.at(d)); };
    for (size_t i = 0; i < graph.edges.size(); i++)
    {
        taken.insert(graph.edges[i].id);
        if (slot.count(graph.edges[i].source) && slot.count(graph.edges[i].dest))
            by_ends.emplace(ends_key(graph.edges[i].source, graph.edges[i].dest), i);
    }
    for (size_t p = 0; p < pairs.size(); p++)
    {
        const size_t source = pairs_by_node_id[p].image_1, dest = pairs_by_node_id[p].image_2;
        const auto existing = by_ends.find(ends_key(source, dest));
        if (existing != by_ends.end())
        {
            graph.edges[existing->second].relations = std::move(relations[p]);
            graph.edges[existing->second].n_relative_poses = 4;
            continue;
        }
        const size_t identifier = graph.draw_edge_id(taken);
        taken.insert(identifier);
        GraphEdge e;
        e.id = identifier;
        e.source = source;
        e.dest = dest;
        e.relations = std::move(relations[p]);
        by_ends.emplace(ends_key(source, dest), graph.edges.size());
        graph.edges.push_back(std::move(e));
        graph.nodes[slot.at(source)].edges.push_back(identifier);
        if (dest != source)
            graph.nodes[slot.at(dest)].edges.push_back(identifier);
    }
    return stats;
}
} // namespace wire
} // namespace ocb_host
