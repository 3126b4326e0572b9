// Stand-in for chrono_dem/utils/ChDemJsonParser.h (reference: src/chrono_dem/utils/ChDemJsonParser.h:33-380): the
// parameter block the Chrono::Dem demos read from a JSON file.  Same struct, field names, accepted keys and enum
// spellings; the JSON reader itself is a small recursive-descent parser (the reference uses rapidjson, which is not
// vendored here): objects, arrays, strings, numbers, true / false / null; only top-level scalar members are used.
#ifndef CHRONO_B200_CHDEMJSONPARSER_H
#define CHRONO_B200_CHDEMJSONPARSER_H
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

#include "chrono_dem/ChDemDefines.h"

namespace chrono {
namespace dem {

/// Structure with Chrono::Dem simulation parameters.
struct ChDemSimulationParameters {
    float sphere_radius;
    float sphere_density;
    float box_X;
    float box_Y;
    float box_Z;
    CHDEM_TIME_INTEGRATOR time_integrator;
    float step_size;
    float time_end;
    float grav_X;
    float grav_Y;
    float grav_Z;
    double normalStiffS2S;
    double normalStiffS2W;
    double normalStiffS2M;
    double normalDampS2S;
    double normalDampS2W;
    double normalDampS2M;
    double tangentDampS2S;
    double tangentDampS2W;
    double tangentDampS2M;
    double tangentStiffS2S;
    double tangentStiffS2W;
    double tangentStiffS2M;
    CHDEM_FRICTION_MODE friction_mode;
    float static_friction_coeffS2S;
    float static_friction_coeffS2W;
    float static_friction_coeffS2M;
    CHDEM_ROLLING_MODE rolling_mode;
    float rolling_friction_coeffS2S;
    float rolling_friction_coeffS2W;
    float rolling_friction_coeffS2M;
    float cohesion_ratio;
    float adhesion_ratio_s2w;
    float adhesion_ratio_s2m;
    CHDEM_VERBOSITY verbose;
    CHDEM_RUN_MODE run_mode;
    unsigned int psi_T;
    unsigned int psi_L;
    float psi_R;
    std::string output_dir;
    std::string checkpoint_file;
    CHDEM_OUTPUT_MODE write_mode;
};

namespace json_detail {

struct Value {
    enum Kind { NUMBER, STRING, BOOL, OTHER } kind = OTHER;
    double num = 0;
    bool is_int = false;
    std::string str;
    bool b = false;
};

class Reader {
  public:
    explicit Reader(const std::string& text) : s(text), i(0) {}
    /// Parses the document; fills `top` with the scalar members of the top-level object.  False on a syntax error.
    bool parse(std::map<std::string, Value>& top) {
        ws();
        if (!eat('{')) return false;
        ws();
        if (eat('}')) return true;
        for (;;) {
            ws();
            std::string key;
            if (!string(key)) return false;
            ws();
            if (!eat(':')) return false;
            Value v;
            if (!value(v)) return false;
            top[key] = v;
            ws();
            if (eat(',')) continue;
            if (eat('}')) break;
            return false;
        }
        ws();
        return i == s.size();
    }

  private:
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) i++; }
    bool eat(char c) { if (i < s.size() && s[i] == c) { i++; return true; } return false; }
    bool string(std::string& out) {
        if (!eat('"')) return false;
        out.clear();
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                const char e = s[i + 1];
                out += (e == 'n') ? '\n' : (e == 't') ? '\t' : e;
                i += 2;
            } else {
                out += s[i++];
            }
        }
        return eat('"');
    }
    bool value(Value& v) {
        ws();
        if (i >= s.size()) return false;
        const char c = s[i];
        if (c == '"') { v.kind = Value::STRING; return string(v.str); }
        if (c == '{' || c == '[') return skip_compound();
        if (!s.compare(i, 4, "true")) { v.kind = Value::BOOL; v.b = true; i += 4; return true; }
        if (!s.compare(i, 5, "false")) { v.kind = Value::BOOL; v.b = false; i += 5; return true; }
        if (!s.compare(i, 4, "null")) { i += 4; return true; }
        char* end = nullptr;
        v.num = std::strtod(s.c_str() + i, &end);
        if (end == s.c_str() + i) return false;
        const std::string tok(s.c_str() + i, (size_t)(end - (s.c_str() + i)));
        v.is_int = tok.find_first_of(".eE") == std::string::npos;
        v.kind = Value::NUMBER;
        i = (size_t)(end - s.c_str());
        return true;
    }
    bool skip_compound() {
        const char open = s[i], close = (open == '{') ? '}' : ']';
        i++;
        ws();
        if (eat(close)) return true;
        for (;;) {
            ws();
            if (open == '{') {
                std::string k;
                if (!string(k)) return false;
                ws();
                if (!eat(':')) return false;
            }
            Value v;
            if (!value(v)) return false;
            ws();
            if (eat(',')) continue;
            return eat(close);
        }
    }
    const std::string& s;
    size_t i;
};

}  // namespace json_detail

/// Print scheme for JSON file with simulation settings.
inline void ShowJSONUsage() {
    std::cout << "JSON fields:\nsphere_radius\nsphere_density\nbox_X\nbox_Y\nbox_Z\n"
                 "time_integrator (forward_euler|chung|centered_difference|extended_taylor)\nstep_size\ntime_end\n"
                 "grav_X\ngrav_Y\ngrav_Z\nnormalStiffS2S\nnormalStiffS2W\nnormalStiffS2M\nnormalDampS2S\nnormalDampS2W\n"
                 "normalDampS2M\ntangentStiffS2S\ntangentStiffS2W\ntangentStiffS2M\ntangentDampS2S\ntangentDampS2W\n"
                 "tangentDampS2M\nfriction_mode (frictionless|single_step|multi_step)\nstatic_friction_coeffS2S\n"
                 "static_friction_coeffS2W\nstatic_friction_coeffS2M\nrolling_mode (no_resistance|schwartz)\n"
                 "rolling_friction_coeffS2S\nrolling_friction_coeffS2W\nrolling_friction_coeffS2M\ncohesion_ratio\n"
                 "adhesion_ratio_s2w\nadhesion_ratio_s2m\nverbose\npsi_T\npsi_L\npsi_R\n"
                 "run_mode (frictionless|one_step|multi_step)\noutput_dir\ncheckpoint_file\nwrite_mode (csv|binary|hdf5|none)"
              << std::endl;
}

/// Flag an invalid arg and print usage.
inline void InvalidArg(const std::string& arg) {
    std::cout << "Invalid arg: " << arg << std::endl;
    ShowJSONUsage();
}

/// Parse the specified JSON file into the params structure.  Members that are absent keep their value; a member of the
/// wrong JSON type is ignored, as in the reference; an unknown enum spelling is an error.
inline bool ParseJSON(const std::string& json_file, ChDemSimulationParameters& params, bool verbose = true) {
    std::ifstream in(json_file);
    if (!in) {
        std::cerr << "Invalid JSON file: " << json_file << std::endl;
        return false;
    }
    std::stringstream buf;
    buf << in.rdbuf();
    const std::string text = buf.str();
    std::map<std::string, json_detail::Value> doc;
    if (!json_detail::Reader(text).parse(doc)) {
        std::cerr << "Invalid JSON file: " << json_file << std::endl;
        return false;
    }
    using json_detail::Value;
    if (verbose)
        std::cout << "--- Parsing JSON ---" << std::endl;
    auto report = [&](const char* k, const std::string& v) {
        if (verbose)
            std::cout << "params." << k << " " << v << std::endl;
    };
    auto num = [&](const char* k, auto& dst) {
        auto it = doc.find(k);
        if (it != doc.end() && it->second.kind == Value::NUMBER) {
            dst = static_cast<std::remove_reference_t<decltype(dst)>>(it->second.num);
            report(k, std::to_string(it->second.num));
        }
    };
    auto integer = [&](const char* k, auto& dst) {
        auto it = doc.find(k);
        if (it != doc.end() && it->second.kind == Value::NUMBER && it->second.is_int) {
            dst = static_cast<std::remove_reference_t<decltype(dst)>>(it->second.num);
            report(k, std::to_string((long long)it->second.num));
        }
    };
    auto str = [&](const char* k, std::string& dst) {
        auto it = doc.find(k);
        if (it != doc.end() && it->second.kind == Value::STRING) {
            dst = it->second.str;
            report(k, dst);
            return true;
        }
        return false;
    };

    num("sphere_radius", params.sphere_radius);
    num("sphere_density", params.sphere_density);
    num("box_X", params.box_X);
    num("box_Y", params.box_Y);
    num("box_Z", params.box_Z);
    std::string e;
    if (str("time_integrator", e)) {
        if (e == "forward_euler") params.time_integrator = CHDEM_TIME_INTEGRATOR::FORWARD_EULER;
        else if (e == "chung") params.time_integrator = CHDEM_TIME_INTEGRATOR::CHUNG;
        else if (e == "centered_difference") params.time_integrator = CHDEM_TIME_INTEGRATOR::CENTERED_DIFFERENCE;
        else if (e == "extended_taylor") params.time_integrator = CHDEM_TIME_INTEGRATOR::EXTENDED_TAYLOR;
        else { InvalidArg("time_integrator"); return false; }
    }
    num("time_end", params.time_end);
    num("grav_X", params.grav_X);
    num("grav_Y", params.grav_Y);
    num("grav_Z", params.grav_Z);
    num("normalStiffS2S", params.normalStiffS2S);
    num("normalStiffS2W", params.normalStiffS2W);
    num("normalStiffS2M", params.normalStiffS2M);
    num("normalDampS2S", params.normalDampS2S);
    num("normalDampS2W", params.normalDampS2W);
    num("normalDampS2M", params.normalDampS2M);
    num("tangentStiffS2S", params.tangentStiffS2S);
    num("tangentStiffS2W", params.tangentStiffS2W);
    num("tangentStiffS2M", params.tangentStiffS2M);
    num("tangentDampS2S", params.tangentDampS2S);
    num("tangentDampS2W", params.tangentDampS2W);
    num("tangentDampS2M", params.tangentDampS2M);
    if (str("friction_mode", e)) {
        if (e == "frictionless") params.friction_mode = CHDEM_FRICTION_MODE::FRICTIONLESS;
        else if (e == "single_step") params.friction_mode = CHDEM_FRICTION_MODE::SINGLE_STEP;
        else if (e == "multi_step") params.friction_mode = CHDEM_FRICTION_MODE::MULTI_STEP;
        else { InvalidArg("friction_mode"); return false; }
    }
    num("static_friction_coeffS2S", params.static_friction_coeffS2S);
    num("static_friction_coeffS2W", params.static_friction_coeffS2W);
    num("static_friction_coeffS2M", params.static_friction_coeffS2M);
    if (str("rolling_mode", e)) {
        if (e == "no_resistance") params.rolling_mode = CHDEM_ROLLING_MODE::NO_RESISTANCE;
        else if (e == "schwartz") params.rolling_mode = CHDEM_ROLLING_MODE::SCHWARTZ;
        else { InvalidArg("rolling_mode"); return false; }
    }
    num("rolling_friction_coeffS2S", params.rolling_friction_coeffS2S);
    num("rolling_friction_coeffS2W", params.rolling_friction_coeffS2W);
    num("rolling_friction_coeffS2M", params.rolling_friction_coeffS2M);
    num("cohesion_ratio", params.cohesion_ratio);
    num("adhesion_ratio_s2w", params.adhesion_ratio_s2w);
    num("adhesion_ratio_s2m", params.adhesion_ratio_s2m);
    {
        int v = -1;
        integer("verbose", v);  // the shipped JSON files say "verbose": false, which is not an int: ignored, as upstream
        if (v >= 0) params.verbose = (CHDEM_VERBOSITY)v;
    }
    integer("psi_T", params.psi_T);
    integer("psi_L", params.psi_L);
    num("psi_R", params.psi_R);
    if (str("run_mode", e)) {
        if (e == "frictionless") params.run_mode = CHDEM_RUN_MODE::FRICTIONLESS;
        else if (e == "one_step") params.run_mode = CHDEM_RUN_MODE::ONE_STEP;
        else if (e == "multi_step") params.run_mode = CHDEM_RUN_MODE::MULTI_STEP;
        else { InvalidArg("run_mode"); return false; }
    }
    str("output_dir", params.output_dir);
    str("checkpoint_file", params.checkpoint_file);
    if (str("write_mode", e)) {
        if (e == "binary") params.write_mode = CHDEM_OUTPUT_MODE::BINARY;
        else if (e == "csv") params.write_mode = CHDEM_OUTPUT_MODE::CSV;
        else if (e == "hdf5") params.write_mode = CHDEM_OUTPUT_MODE::HDF5;
        else if (e == "none") params.write_mode = CHDEM_OUTPUT_MODE::NONE;
        else { InvalidArg("write_mode"); return false; }
    }
    num("step_size", params.step_size);
    if (verbose)
        std::cout << "--------------------" << std::endl;
    return true;
}

}  // namespace dem
}  // namespace chrono
#endif
