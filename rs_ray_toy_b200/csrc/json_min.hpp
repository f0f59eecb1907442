// A small recursive-descent JSON reader for scene.json (the reference uses serde_json 1.0.64,
// src/renderprocess.rs:101; numbers are parsed with strtod, which like serde_json yields the
// nearest f64).  Objects keep insertion order; duplicate keys keep the last value.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace rrt {
namespace json {

struct Value;
using ValuePtr = std::shared_ptr<Value>;
struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> obj;

    const Value* get(const std::string& key) const {
        if (kind != Object) return nullptr;
        const Value* found = nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) found = kv.second.get();
        return found;
    }
    void set(const std::string& key, ValuePtr v) {
        for (auto& kv : obj)
            if (kv.first == key) {
                kv.second = v;
                return;
            }
        obj.emplace_back(key, v);
    }
    bool is_string() const { return kind == String; }
    bool is_array() const { return kind == Array; }
    bool is_object() const { return kind == Object; }
    bool is_number() const { return kind == Number; }
};

class Parser {
  public:
    explicit Parser(const std::string& text) : s_(text) {}
    ValuePtr parse() {
        ValuePtr v = value();
        ws();
        if (i_ != s_.size()) fail("trailing characters");
        return v;
    }

  private:
    const std::string& s_;
    size_t i_ = 0;
    [[noreturn]] void fail(const char* what) const {
        throw std::runtime_error(std::string("JSON: ") + what + " at byte " + std::to_string(i_));
    }
    void ws() {
        while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\t' || s_[i_] == '\n' || s_[i_] == '\r')) ++i_;
    }
    ValuePtr value() {
        ws();
        if (i_ >= s_.size()) fail("unexpected end");
        char c = s_[i_];
        auto v = std::make_shared<Value>();
        if (c == '{') {
            v->kind = Value::Object;
            ++i_;
            ws();
            if (i_ < s_.size() && s_[i_] == '}') {
                ++i_;
                return v;
            }
            for (;;) {
                ws();
                if (i_ >= s_.size() || s_[i_] != '"') fail("expected a key");
                std::string k = string();
                ws();
                if (i_ >= s_.size() || s_[i_] != ':') fail("expected ':'");
                ++i_;
                v->set(k, value());
                ws();
                if (i_ < s_.size() && s_[i_] == ',') {
                    ++i_;
                    continue;
                }
                if (i_ < s_.size() && s_[i_] == '}') {
                    ++i_;
                    return v;
                }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            v->kind = Value::Array;
            ++i_;
            ws();
            if (i_ < s_.size() && s_[i_] == ']') {
                ++i_;
                return v;
            }
            for (;;) {
                v->arr.push_back(value());
                ws();
                if (i_ < s_.size() && s_[i_] == ',') {
                    ++i_;
                    continue;
                }
                if (i_ < s_.size() && s_[i_] == ']') {
                    ++i_;
                    return v;
                }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') {
            v->kind = Value::String;
            v->str = string();
            return v;
        }
        if (s_.compare(i_, 4, "true") == 0) {
            v->kind = Value::Bool;
            v->b = true;
            i_ += 4;
            return v;
        }
        if (s_.compare(i_, 5, "false") == 0) {
            v->kind = Value::Bool;
            i_ += 5;
            return v;
        }
        if (s_.compare(i_, 4, "null") == 0) {
            i_ += 4;
            return v;
        }
        const char* start = s_.c_str() + i_;
        char* end = nullptr;
        double d = std::strtod(start, &end);
        if (end == start) fail("unexpected character");
        v->kind = Value::Number;
        v->num = d;
        i_ += (size_t)(end - start);
        return v;
    }
    std::string string() {
        std::string out;
        ++i_;  // opening quote
        while (i_ < s_.size() && s_[i_] != '"') {
            char c = s_[i_++];
            if (c == '\\') {
                if (i_ >= s_.size()) fail("bad escape");
                char e = s_[i_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (i_ + 4 > s_.size()) fail("bad \\u escape");
                        unsigned cp = (unsigned)std::strtoul(s_.substr(i_, 4).c_str(), nullptr, 16);
                        i_ += 4;
                        if (cp < 0x80) {
                            out += (char)cp;
                        } else if (cp < 0x800) {
                            out += (char)(0xC0 | (cp >> 6));
                            out += (char)(0x80 | (cp & 0x3F));
                        } else {
                            out += (char)(0xE0 | (cp >> 12));
                            out += (char)(0x80 | ((cp >> 6) & 0x3F));
                            out += (char)(0x80 | (cp & 0x3F));
                        }
                        break;
                    }
                    default: out += e;
                }
            } else {
                out += c;
            }
        }
        if (i_ >= s_.size()) fail("unterminated string");
        ++i_;
        return out;
    }
};

inline ValuePtr parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace json
}  // namespace rrt
