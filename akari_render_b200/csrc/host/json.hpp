// Minimal JSON reader for the scenegraph / method files (no third-party deps).
// Objects keep their members in a std::map so iteration is in byte-wise key order, which is what
// the reference's `Collection<T>(BTreeMap<NodeRef<T>, T>)` does
// (reference: crates/akari_scenegraph/src/lib.rs:71) — ids are assigned in that order.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace akr::json {

struct Value;
using Object = std::map<std::string, Value>;
using Array = std::vector<Value>;

struct Value {
    enum Type { Null, Bool, Number, String, ArrayT, ObjectT } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::shared_ptr<Array> arr;
    std::shared_ptr<Object> obj;

    bool is_null() const { return type == Null; }
    bool is_object() const { return type == ObjectT; }
    bool is_array() const { return type == ArrayT; }
    bool is_string() const { return type == String; }
    bool is_number() const { return type == Number; }

    const Object &object() const {
        if (type != ObjectT) throw std::runtime_error("json: expected object");
        return *obj;
    }
    const Array &array() const {
        if (type != ArrayT) throw std::runtime_error("json: expected array");
        return *arr;
    }
    const std::string &string() const {
        if (type != String) throw std::runtime_error("json: expected string");
        return str;
    }
    double number() const {
        if (type != Number) throw std::runtime_error("json: expected number");
        return num;
    }
    bool boolean() const {
        if (type != Bool) throw std::runtime_error("json: expected bool");
        return b;
    }
    bool has(const std::string &k) const { return type == ObjectT && obj->count(k) != 0; }
    const Value &at(const std::string &k) const {
        const Object &o = object();
        auto it = o.find(k);
        if (it == o.end()) throw std::runtime_error("json: missing key '" + k + "'");
        return it->second;
    }
    const Value &at(size_t i) const {
        const Array &a = array();
        if (i >= a.size()) throw std::runtime_error("json: index out of range");
        return a[i];
    }
    // serde deserialises JSON numbers into f32 fields by parsing to f64 and casting.
    float f32() const { return static_cast<float>(number()); }
    uint32_t u32() const { return static_cast<uint32_t>(number()); }
    int32_t i32() const { return static_cast<int32_t>(number()); }
    uint64_t u64() const { return static_cast<uint64_t>(number()); }
};

class Parser {
  public:
    explicit Parser(const std::string &text) : s_(text.data()), n_(text.size()) {}
    Value parse() {
        Value v = value();
        ws();
        if (i_ != n_) fail("trailing characters");
        return v;
    }

  private:
    const char *s_;
    size_t n_, i_ = 0;
    int depth_ = 0;  // nesting of arrays / objects: bounded like serde_json's recursion limit (128), the parser is recursive
    struct Nest {
        Parser &p;
        explicit Nest(Parser &parser) : p(parser) {
            if (++p.depth_ > 128) p.fail("recursion limit exceeded");
        }
        ~Nest() { --p.depth_; }
    };
    [[noreturn]] void fail(const char *msg) {
        throw std::runtime_error(std::string("json: ") + msg + " at byte " + std::to_string(i_));
    }
    void ws() {
        while (i_ < n_ && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\r' || s_[i_] == '\t')) ++i_;
    }
    Value value() {
        ws();
        if (i_ >= n_) fail("unexpected end");
        char c = s_[i_];
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') {
            Value v;
            v.type = Value::String;
            v.str = string();
            return v;
        }
        if (c == 't' && n_ - i_ >= 4 && !std::memcmp(s_ + i_, "true", 4)) {
            i_ += 4;
            Value v;
            v.type = Value::Bool;
            v.b = true;
            return v;
        }
        if (c == 'f' && n_ - i_ >= 5 && !std::memcmp(s_ + i_, "false", 5)) {
            i_ += 5;
            Value v;
            v.type = Value::Bool;
            v.b = false;
            return v;
        }
        if (c == 'n' && n_ - i_ >= 4 && !std::memcmp(s_ + i_, "null", 4)) {
            i_ += 4;
            return Value{};
        }
        return number();
    }
    Value number() {
        const char *start = s_ + i_;
        char *end = nullptr;
        double d = std::strtod(start, &end);
        if (end == start) fail("bad number");
        i_ += static_cast<size_t>(end - start);
        Value v;
        v.type = Value::Number;
        v.num = d;
        return v;
    }
    std::string string() {
        ++i_;  // opening quote
        std::string out;
        while (true) {
            if (i_ >= n_) fail("unterminated string");
            char c = s_[i_++];
            if (c == '"') break;
            if (c == '\\') {
                if (i_ >= n_) fail("bad escape");
                char e = s_[i_++];
                switch (e) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    if (n_ - i_ < 4) fail("bad \\u escape");
                    unsigned cp = static_cast<unsigned>(std::strtoul(std::string(s_ + i_, 4).c_str(), nullptr, 16));
                    i_ += 4;
                    if (cp < 0x80) {
                        out.push_back(static_cast<char>(cp));
                    } else if (cp < 0x800) {
                        out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
                        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
                    } else {
                        out.push_back(static_cast<char>(0xE0 | (cp >> 12)));
                        out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
                        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
                    }
                    break;
                }
                default: fail("bad escape");
                }
            } else {
                out.push_back(c);
            }
        }
        return out;
    }
    Value array() {
        Nest nest(*this);
        ++i_;
        Value v;
        v.type = Value::ArrayT;
        v.arr = std::make_shared<Array>();
        ws();
        if (i_ < n_ && s_[i_] == ']') {
            ++i_;
            return v;
        }
        while (true) {
            v.arr->push_back(value());
            ws();
            if (i_ >= n_) fail("unterminated array");
            if (s_[i_] == ',') {
                ++i_;
                continue;
            }
            if (s_[i_] == ']') {
                ++i_;
                break;
            }
            fail("expected , or ]");
        }
        return v;
    }
    Value object() {
        Nest nest(*this);
        ++i_;
        Value v;
        v.type = Value::ObjectT;
        v.obj = std::make_shared<Object>();
        ws();
        if (i_ < n_ && s_[i_] == '}') {
            ++i_;
            return v;
        }
        while (true) {
            ws();
            if (i_ >= n_ || s_[i_] != '"') fail("expected key");
            std::string k = string();
            ws();
            if (i_ >= n_ || s_[i_] != ':') fail("expected :");
            ++i_;
            (*v.obj)[k] = value();
            ws();
            if (i_ >= n_) fail("unterminated object");
            if (s_[i_] == ',') {
                ++i_;
                continue;
            }
            if (s_[i_] == '}') {
                ++i_;
                break;
            }
            fail("expected , or }");
        }
        return v;
    }
};

inline Value parse(const std::string &text) { return Parser(text).parse(); }

}  // namespace akr::json
