// ArgList: the command-line parser of the Channelflow programs (reference cfbasics/arglist.h:27-110): named options with
// short and long forms, flags, positional arguments counted from the END of the line, -h/--help listing, unused-argument
// check, and a record of the command line in <program>.args.
#ifndef CFB200_ARGLIST_H
#define CFB200_ARGLIST_H
#include <unistd.h>

#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "cfbasics/cfbasics.h"
#include "cfbasics/mathdefs.h"

namespace chflow {

// a real argument is a number or the name of a file holding one
inline Real arg2real(const std::string& s) {
    Real r = 0;
    if (fileExists(s)) load(r, s);
    else r = std::atof(s.c_str());
    return r;
}

class ArgList {
   public:
    typedef std::string str;
    ArgList() {}
    ArgList(int argc, char* argv[], const str& purpose) : args_(argv, argv + argc), used_(argc, false) {
        for (int i = 0; i < argc; ++i)
            if (args_[i] == "-h" || args_[i] == "--help") { helpmode_ = true; used_[i] = true; }
        if (helpmode_) std::cerr << argv[0] << " : \n\t" << purpose << std::endl << std::endl;
        if (argc > 0) used_[0] = true;
    }
    bool helpmode() const { return helpmode_; }
    bool errormode() const { return errormode_; }
    int remaining() const { int n = 0; for (bool u : used_) n += u ? 0 : 1; return n; }
    void section(const str& name, const str& description = "") const {
        if (helpmode_) std::cerr << '\n' << name << (description.empty() ? "" : " : " + description) << '\n';
    }

    bool getflag(const str& s, const str& l, const str& help) {
        if (helpmode_) { printhelp(s, l, "", "", help); return false; }
        const int i = find(s, l);
        if (i < 0) return false;
        used_[i] = true;
        return true;
    }
    bool getbool(const str& s, const str& l, const str& help) { return tobool(required(s, l, "bool", help)); }
    int getint(const str& s, const str& l, const str& help) { return std::atoi(required(s, l, "int", help).c_str()); }
    Real getreal(const str& s, const str& l, const str& help) { return arg2real(required(s, l, "real", help)); }
    str getstr(const str& s, const str& l, const str& help) { return required(s, l, "string", help); }
    str getpath(const str& s, const str& l, const str& help) { return pathfix(required(s, l, "path", help)); }

    bool getbool(const str& s, const str& l, bool d, const str& help) {
        str v;
        return optional(s, l, "bool", d ? "true" : "false", help, v) ? tobool(v) : d;
    }
    int getint(const str& s, const str& l, int d, const str& help) {
        str v;
        return optional(s, l, "int", std::to_string(d), help, v) ? std::atoi(v.c_str()) : d;
    }
    Real getreal(const str& s, const str& l, Real d, const str& help) {
        str v;
        return optional(s, l, "real", r2s(d), help, v) ? arg2real(v) : d;
    }
    str getstr(const str& s, const str& l, const str& d, const str& help) {
        str v;
        return optional(s, l, "string", d, help, v) ? v : d;
    }
    str getpath(const str& s, const str& l, const str& d, const str& help) {
        str v;
        return pathfix(optional(s, l, "path", d, help, v) ? v : d);
    }

    // positional arguments: position 1 is the LAST argument of the line
    str getstr(int position, const str& meaning, const str& help) {
        if (helpmode_) { printhelp(position, meaning, help); return ""; }
        const int i = (int)args_.size() - position;
        if (i < 1 || used_[i]) {
            std::cerr << "error : missing or already-used positional argument " << meaning << " (position " << position << " from the end)\n";
            errormode_ = true;
            return "";
        }
        used_[i] = true;
        return args_[i];
    }
    str getpath(int position, const str& meaning, const str& help) { return pathfix(getstr(position, meaning, help)); }
    Real getreal(int position, const str& meaning, const str& help) { return arg2real(getstr(position, meaning, help)); }
    std::vector<std::string> remainingatend() {
        std::vector<std::string> r;
        int i = (int)args_.size() - 1;
        while (i >= 1 && !used_[i]) --i;
        for (int j = i + 1; j < (int)args_.size(); ++j) { r.push_back(args_[j]); used_[j] = true; }
        return r;
    }

    void save(const str& outdir) const {
        if (mpirank() != 0 || args_.empty()) return;
        str prog = args_[0];
        const size_t s = prog.find_last_of('/');
        if (s != str::npos) prog = prog.substr(s + 1);
        std::ofstream os((pathfix(outdir) + prog + ".args").c_str(), std::ios::app);
        for (const str& a : args_) os << a << ' ';
        os << '\n';
    }
    void save() const { save("./"); }
    // stop on -h (after the listing), on errors, and on arguments nobody asked for
    void check() {
        if (helpmode_) std::exit(0);
        bool stray = false;
        for (size_t i = 1; i < args_.size(); ++i)
            if (!used_[i]) { std::cerr << "error : unrecognized option/value " << args_[i] << std::endl; stray = true; }
        if (stray || errormode_) {
            std::cerr << "Please rerun the program with option -h or --help for a listing of valid options" << std::endl;
            std::exit(1);
        }
    }

   private:
    std::vector<str> args_;
    std::vector<bool> used_;
    bool helpmode_ = false, errormode_ = false;

    int find(const str& s, const str& l) const {
        for (size_t i = 1; i < args_.size(); ++i)
            if (!used_[i] && (args_[i] == s || args_[i] == l)) return (int)i;
        return -1;
    }
    static bool tobool(const str& v) { return v == "true" || v == "1" || v == "True" || v == "TRUE" || v == "t"; }
    bool optional(const str& s, const str& l, const str& type, const str& dflt, const str& help, str& value) {
        if (helpmode_) { printhelp(s, l, type, dflt, help); return false; }
        const int i = find(s, l);
        if (i < 0) return false;
        if (i + 1 >= (int)args_.size()) {
            std::cerr << "error : option " << args_[i] << " should be followed by a value of type " << type << '\n';
            errormode_ = true;
            used_[i] = true;
            return false;
        }
        used_[i] = used_[i + 1] = true;
        value = args_[i + 1];
        return true;
    }
    str required(const str& s, const str& l, const str& type, const str& help) {
        str v;
        if (helpmode_) { printhelp(s, l, type, "", help); return ""; }
        if (!optional(s, l, type, "", help, v) && !errormode_) {
            std::cerr << "error : required option " << s << " or " << l << " <" << type << "> is missing (" << help << ")\n";
            errormode_ = true;
        }
        return v;
    }
    void printhelp(const str& s, const str& l, const str& type, const str& dflt, const str& help) {
        std::cerr << "  " << std::setw(8) << std::left << s << std::setw(20) << l << std::setw(10) << (type.empty() ? "" : "<" + type + ">");
        if (!dflt.empty()) std::cerr << "default == " << std::setw(12) << dflt;
        std::cerr << help << std::right << std::endl;
    }
    void printhelp(int position, const str& name, const str& help) {
        std::cerr << "  " << name << "  (trailing arg " << position << ")  " << help << std::endl;
    }
};

}  // namespace chflow
#endif
