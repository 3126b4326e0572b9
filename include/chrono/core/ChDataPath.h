// Minimal stand-in for chrono/core/ChDataPath.h (reference: src/chrono/core/ChDataPath.h): where demos look for their
// input data and put their output.  Defaults can be overridden with the CHRONO_DATA_DIR / CHRONO_OUTPUT_DIR environment
// variables (additive) or SetChronoDataPath / SetChronoOutputPath (reference API).
#ifndef CHRONO_B200_CHDATAPATH_H
#define CHRONO_B200_CHDATAPATH_H
#include <cstdlib>
#include <filesystem>
#include <string>

namespace chrono {

namespace detail {
inline std::string& data_path() {
    static std::string p = [] {
        const char* e = std::getenv("CHRONO_DATA_DIR");
        std::string s = e ? e : "../data/";
        if (!s.empty() && s.back() != '/') s += '/';
        return s;
    }();
    return p;
}
inline std::string& output_path() {
    static std::string p = [] {
        const char* e = std::getenv("CHRONO_OUTPUT_DIR");
        std::string s = e ? e : "DEMO_OUTPUT/";
        if (!s.empty() && s.back() != '/') s += '/';
        return s;
    }();
    return p;
}
}  // namespace detail

inline void SetChronoDataPath(const std::string& path) { detail::data_path() = path; }
inline const std::string& GetChronoDataPath() { return detail::data_path(); }
inline std::string GetChronoDataFile(const std::string& filename) { return detail::data_path() + filename; }
inline void SetChronoOutputPath(const std::string& path) { detail::output_path() = path; }
inline const std::string& GetChronoOutputPath() { return detail::output_path(); }
/// Create the directory (and its parents) if it does not exist yet; false if that is impossible.
inline bool CreateOutputDirectory(const std::filesystem::path& dir) {
    std::error_code ec;
    if (std::filesystem::exists(dir, ec))
        return std::filesystem::is_directory(dir, ec);
    return std::filesystem::create_directories(dir, ec) || std::filesystem::is_directory(dir, ec);
}

}  // namespace chrono
#endif
