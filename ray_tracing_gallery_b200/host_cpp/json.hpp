// json.hpp — the little JSON a GLB header needs (objects, arrays, strings, numbers, true/false/null).
// Part of the C++ host side of libb200rt (the reference's host is compiled Rust; see host.hpp).
#pragma once
#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace b200rt_host {

struct Json {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    const Json* find(const std::string& key) const {
        if (type != Obj) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const std::string& key) const { return find(key) != nullptr; }
    const Json& at(const std::string& key) const {
        const Json* j = find(key);
        if (!j) throw std::runtime_error("glTF JSON: missing key '" + key + "'");
        return *j;
    }
    const Json& at(size_t i) const {
        if (type != Arr || i >= arr.size()) throw std::runtime_error("glTF JSON: array index out of range");
        return arr[i];
    }
    size_t size() const { return type == Arr ? arr.size() : type == Obj ? obj.size() : 0; }
    double number() const {
        if (type != Num) throw std::runtime_error("glTF JSON: number expected");
        return num;
    }
    long long integer() const { return (long long)number(); }
    const std::string& string() const {
        if (type != Str) throw std::runtime_error("glTF JSON: string expected");
        return str;
    }
    double number_or(const std::string& key, double def) const {
        const Json* j = find(key);
        return j ? j->number() : def;
    }
    long long integer_or(const std::string& key, long long def) const {
        const Json* j = find(key);
        return j ? j->integer() : def;
    }
    std::string string_or(const std::string& key, const std::string& def) const {
        const Json* j = find(key);
        return j ? j->string() : def;
    }
};

class JsonParser {
public:
    explicit JsonParser(const std::string& s) : s_(s) {}
    Json parse() {
        Json j = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return j;
    }

private:
    const std::string& s_;
    size_t p_ = 0;
    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("glTF JSON: ") + what + " at byte " + std::to_string(p_)); }
    void ws() {
        while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) p_++;
    }
    bool eat(char c) {
        ws();
        if (p_ < s_.size() && s_[p_] == c) { p_++; return true; }
        return false;
    }
    void expect(char c) {
        if (!eat(c)) fail("unexpected character");
    }
    static void utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    unsigned hex4() {
        if (p_ + 4 > s_.size()) fail("bad \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; i++) {
            char c = s_[p_++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string string_body() {
        std::string out;
        for (;;) {
            if (p_ >= s_.size()) fail("unterminated string");
            char c = s_[p_++];
            if (c == '"') return out;
            if (c != '\\') { out += c; continue; }
            if (p_ >= s_.size()) fail("bad escape");
            char e = s_[p_++];
            switch (e) {
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                case '/': out += '/'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;
                case 't': out += '\t'; break;
                case 'u': {
                    unsigned cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && p_ + 1 < s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
                        p_ += 2;
                        unsigned lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    utf8(out, cp);
                    break;
                }
                default: fail("bad escape");
            }
        }
    }
    Json value() {
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        char c = s_[p_];
        Json j;
        if (c == '{') {
            p_++;
            j.type = Json::Obj;
            if (eat('}')) return j;
            do {
                ws();
                expect('"');
                std::string k = string_body();
                expect(':');
                j.obj.emplace_back(std::move(k), value());
            } while (eat(','));
            expect('}');
        } else if (c == '[') {
            p_++;
            j.type = Json::Arr;
            if (eat(']')) return j;
            do j.arr.push_back(value());
            while (eat(','));
            expect(']');
        } else if (c == '"') {
            p_++;
            j.type = Json::Str;
            j.str = string_body();
        } else if (s_.compare(p_, 4, "true") == 0) {
            p_ += 4; j.type = Json::Bool; j.b = true;
        } else if (s_.compare(p_, 5, "false") == 0) {
            p_ += 5; j.type = Json::Bool; j.b = false;
        } else if (s_.compare(p_, 4, "null") == 0) {
            p_ += 4;
        } else {
            const char* start = s_.c_str() + p_;
            char* end = nullptr;
            double v = std::strtod(start, &end);
            if (end == start) fail("bad value");
            p_ += (size_t)(end - start);
            j.type = Json::Num;
            j.num = v;
        }
        return j;
    }
};

inline Json parse_json(const std::string& text) { return JsonParser(text).parse(); }

}  // namespace b200rt_host
