// Minimal stand-in for the subset of cxxopts v3 that the reference's apps use (examples/multiply.cpp,
// examples/cublasXt-multiply.cpp, tests/test-multiply.cpp): Options(name, description).add_options()(spec, help, value<T>()
// ->default_value(text)), parse(argc, argv), result[name].as<T>(), result.count(name), help().
// cxxopts itself is fetched from GitHub by the reference's CMake (CMakeLists.txt:65-72) and is not in this image; with this
// header on the include path the reference's apps compile UNCHANGED against include/Tiled-MM (tools/build_ref_apps.sh).
// Written from the documented interface; shares no code with cxxopts.  Not part of the product library.
#pragma once
// the real header pulls these in and the reference apps rely on that (std::unordered_set, std::transform, std::toupper)
#include <algorithm>
#include <cctype>
#include <unordered_map>
#include <unordered_set>

#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cxxopts {

class Value {
public:
    std::shared_ptr<Value> default_value(const std::string& text) {
        has_default_ = true;
        default_ = text;
        return self_.lock();
    }
    bool has_default() const { return has_default_; }
    const std::string& default_text() const { return default_; }
    bool is_flag = false;
    std::weak_ptr<Value> self_;

private:
    bool has_default_ = false;
    std::string default_;
};

template <typename T>
std::shared_ptr<Value> value() {
    auto v = std::make_shared<Value>();
    v->self_ = v;
    v->is_flag = std::is_same<T, bool>::value;
    return v;
}

struct Spec {
    std::string short_name, long_name, help;
    std::shared_ptr<Value> value;
};

class OptionValue {
public:
    OptionValue() = default;
    OptionValue(std::string text, bool present) : text_(std::move(text)), present_(present) {}
    template <typename T>
    T as() const {
        if (!present_) throw std::runtime_error("option has no value");
        return convert<T>(text_);
    }

private:
    template <typename T>
    static T convert(const std::string& s) {
        std::istringstream in(s);
        T out{};
        in >> out;
        if (in.fail()) throw std::runtime_error("cannot parse option value '" + s + "'");
        return out;
    }
    std::string text_;
    bool present_ = false;
};
template <>
inline std::string OptionValue::convert<std::string>(const std::string& s) { return s; }
template <>
inline bool OptionValue::convert<bool>(const std::string& s) { return !(s.empty() || s == "0" || s == "false"); }

class ParseResult {
public:
    OptionValue operator[](const std::string& name) const {
        auto it = values_.find(name);
        if (it == values_.end()) throw std::runtime_error("option '" + name + "' is not declared");
        return it->second;
    }
    std::size_t count(const std::string& name) const {
        auto it = counts_.find(name);
        return it == counts_.end() ? 0 : it->second;
    }
    std::map<std::string, OptionValue> values_;
    std::map<std::string, std::size_t> counts_;
};

class Options;
class OptionAdder {
public:
    explicit OptionAdder(Options& o) : owner_(o) {}
    OptionAdder& operator()(const std::string& spec, const std::string& help, std::shared_ptr<Value> v = value<bool>());

private:
    Options& owner_;
};

class Options {
public:
    Options(std::string program, std::string description = "") : program_(std::move(program)), description_(std::move(description)) {
        specs_.push_back({"h", "help", "Print usage", value<bool>()});
    }
    OptionAdder add_options(const std::string& = "") { return OptionAdder(*this); }

    ParseResult parse(int argc, const char* const* argv) const {
        ParseResult r;
        for (const Spec& s : specs_)
            if (s.value->has_default()) r.values_[s.long_name] = OptionValue(s.value->default_text(), true);
        for (int i = 1; i < argc; ++i) {
            std::string arg = argv[i], name, inline_value;
            bool has_inline = false;
            const Spec* spec = nullptr;
            if (arg.rfind("--", 0) == 0) {
                name = arg.substr(2);
                const auto eq = name.find('=');
                if (eq != std::string::npos) { inline_value = name.substr(eq + 1); name = name.substr(0, eq); has_inline = true; }
                spec = find_long(name);
            } else if (arg.size() >= 2 && arg[0] == '-') {
                name = arg.substr(1, 1);
                if (arg.size() > 2) { inline_value = arg.substr(2); has_inline = true; }
                spec = find_short(name);
            }
            if (!spec) throw std::runtime_error("Option '" + arg + "' does not exist");
            std::string text = "true";
            if (!spec->value->is_flag) {
                if (has_inline) text = inline_value;
                else if (i + 1 < argc) text = argv[++i];
                else throw std::runtime_error("Option '" + arg + "' is missing an argument");
            }
            r.values_[spec->long_name] = OptionValue(text, true);
            r.counts_[spec->long_name]++;
        }
        return r;
    }

    std::string help() const {
        std::ostringstream o;
        o << description_ << "\nUsage:\n  " << program_ << " [OPTION...]\n\n";
        for (const Spec& s : specs_) {
            std::string left = "  " + (s.short_name.empty() ? std::string("    ") : "-" + s.short_name + ", ") + "--" + s.long_name + (s.value->is_flag ? "" : " arg");
            if (left.size() < 28) left.resize(28, ' ');
            o << left << " " << s.help;
            if (s.value->has_default()) o << " (default: " << s.value->default_text() << ")";
            o << "\n";
        }
        return o.str();
    }

    void add(const std::string& spec, const std::string& help, std::shared_ptr<Value> v) {
        Spec s;
        const auto comma = spec.find(',');
        if (comma == std::string::npos) s.long_name = spec;
        else { s.short_name = spec.substr(0, comma); s.long_name = spec.substr(comma + 1); }
        s.help = help;
        s.value = std::move(v);
        specs_.push_back(std::move(s));
    }

private:
    const Spec* find_long(const std::string& n) const {
        for (const Spec& s : specs_) if (s.long_name == n) return &s;
        return nullptr;
    }
    const Spec* find_short(const std::string& n) const {
        for (const Spec& s : specs_) if (!s.short_name.empty() && s.short_name == n) return &s;
        return nullptr;
    }
    std::string program_, description_;
    std::vector<Spec> specs_;
};

inline OptionAdder& OptionAdder::operator()(const std::string& spec, const std::string& help, std::shared_ptr<Value> v) {
    owner_.add(spec, help, std::move(v));
    return *this;
}

}  // namespace cxxopts
