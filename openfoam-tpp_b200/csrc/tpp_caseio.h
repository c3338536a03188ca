// tpp_caseio.h - OpenFOAM case directories read and written inside the library (host C++ only).
//
// SURVEY.md 8(b) sketches the boundary as `tpp_open(case_dir)`: what `foamRun` does when
// `make run` / `make resume` start it in a case directory
// (/root/reference/circularSloshingTank/Makefile:71-99, /root/reference/main.py:333-348).  The Python
// host of this repo has its own reader (openfoam_tpp_b200/foamfile.py, case.py); this file is the
// same contract for hosts that are not Python: they bind tpp_open / tpp_run_case and need no
// FoamFile code of their own.  Files read:
//   constant/polyMesh/{points,faces,owner,neighbour,boundary,cellZones}   (ascii or binary, gmshToFoam layout)
//   <start time>/{alpha.water,U,p_rgh[,phi,Uf]} [+ uniform/time]          (0/ after setFields, or a restart)
//   system/{controlDict,fvSchemes,fvSolution[,functions]}, constant/{g,momentumTransport,phaseProperties,
//   physicalProperties.water,physicalProperties.air[,dynamicMeshDict + its 6DoF table]}
// Keywords this solver cannot honour are errors that name the file and the keyword; nothing is
// defaulted silently (same rules as case.read_config).
#pragma once
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <regex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/tppvof.h"

namespace caseio {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
[[noreturn]] inline void fail(const std::string& m) { throw Error(m); }

inline bool exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}
inline bool isDir(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
inline void makeDirs(const std::string& p) {
    for (size_t i = 1; i <= p.size(); i++)
        if (i == p.size() || p[i] == '/') {
            std::string q = p.substr(0, i);
            if (!isDir(q) && mkdir(q.c_str(), 0777) != 0 && !isDir(q)) fail(q + ": cannot create directory");
        }
}
inline std::string slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) fail(path + ": cannot open");
    std::string s;
    struct stat st;
    if (fstat(fileno(f), &st) == 0 && st.st_size > 0) s.resize((size_t)st.st_size);  // one read of the whole file
    size_t got = s.empty() ? 0 : fread(&s[0], 1, s.size(), f);
    s.resize(got);
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);  // whatever a growing file still holds
    fclose(f);
    return s;
}

// ----------------------------------------------------------------------------------------
// value tree of a dictionary file
// ----------------------------------------------------------------------------------------
struct Node;
typedef std::vector<Node> Stream;                          // the tokens of one entry up to ';'
typedef std::vector<std::pair<std::string, Stream>> Dict;  // entries in file order
struct Node {
    enum Kind { WORD, LIST, DICT, NUM } kind = WORD;
    std::string w;          // WORD: the token;  DICT inside a list: its name (`name { ... }`)
    std::vector<Node> l;    // LIST
    Dict d;                 // DICT
    std::vector<double> a;  // NUM: n x nc numbers of a `List<T> N (...)`
    int nc = 1;
    long n = 0;
};

inline const Stream* find(const Dict& d, const std::string& key) {
    for (auto& kv : d)
        if (kv.first == key) return &kv.second;
    const Stream* hit = nullptr;  // OpenFOAM regular-expression keys ("pcorr.*"): the last match wins
    for (auto& kv : d)
        if (kv.first.size() > 2 && kv.first[0] == '"') {
            try {
                if (std::regex_match(key, std::regex(kv.first.substr(1, kv.first.size() - 2)))) hit = &kv.second;
            } catch (const std::regex_error&) {
            }
        }
    return hit;
}
inline const Stream& lookup(const Dict& d, const std::string& key, const std::string& path) {
    const Stream* s = find(d, key);
    if (!s) fail("keyword '" + key + "' is missing in " + path);
    return *s;
}
inline const Dict& subDict(const Dict& d, const std::string& key, const std::string& path) {
    const Stream& s = lookup(d, key, path);
    if (s.size() != 1 || s[0].kind != Node::DICT) fail(path + ": '" + key + "' is not a dictionary");
    return s[0].d;
}
inline std::string join(const Stream& s) {
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        if (i) o += " ";
        const Node& v = s[i];
        if (v.kind == Node::WORD) o += v.w;
        else if (v.kind == Node::LIST) o += "(" + join(v.l) + ")";
        else if (v.kind == Node::NUM) {
            o += std::to_string(v.n) + "(";
            char b[40];
            for (size_t k = 0; k < v.a.size(); k++) {
                snprintf(b, sizeof b, "%.17g", v.a[k]);
                o += (k ? " " : "") + std::string(b);
            }
            o += ")";
        } else o += "{...}";
    }
    return o;
}
inline std::string word(const Dict& d, const std::string& key, const std::string& path) { return join(lookup(d, key, path)); }
inline std::string wordOr(const Dict& d, const std::string& key, const std::string& def) {
    const Stream* s = find(d, key);
    return s ? join(*s) : def;
}
inline double toNumber(const std::string& t, const std::string& what) {
    char* e = nullptr;
    double v = strtod(t.c_str(), &e);
    if (t.empty() || e == t.c_str() || *e) fail("expected a number for " + what + ", got '" + t + "'");
    return v;
}
// `[dims] value`, `uniform 0` or a bare number: the last token counts
inline double number(const Stream& s, const std::string& what) {
    if (s.empty() || s.back().kind != Node::WORD) fail("expected a number for " + what + ", got '" + join(s) + "'");
    return toNumber(s.back().w, what);
}
inline double number(const Dict& d, const std::string& key, const std::string& path) { return number(lookup(d, key, path), path + ":" + key); }
inline double numberOr(const Dict& d, const std::string& key, const std::string& path, double def) {
    const Stream* s = find(d, key);
    return s ? number(*s, path + ":" + key) : def;
}
inline int integerOr(const Dict& d, const std::string& key, const std::string& path, int def) {
    const Stream* s = find(d, key);
    if (!s) return def;
    double v = number(*s, path + ":" + key);
    if (v != std::floor(v)) fail("expected an integer for " + path + ":" + key + ", got '" + join(*s) + "'");
    return (int)v;
}
inline void vector3(const Stream& s, const std::string& what, double* out) {
    const std::vector<Node>* v = nullptr;
    if (s.size() == 3 && s[0].kind == Node::WORD) v = &s;  // bare `x y z`
    else if (!s.empty() && s.back().kind == Node::LIST) v = &s.back().l;
    if (!v || v->size() != 3) fail("expected a vector for " + what + ", got '" + join(s) + "'");
    for (int k = 0; k < 3; k++) {
        if ((*v)[k].kind != Node::WORD) fail("expected a vector for " + what + ", got '" + join(s) + "'");
        out[k] = toNumber((*v)[k].w, what);
    }
}
inline bool toSwitch(std::string s, const std::string& what) {
    std::string t = s;
    for (auto& c : t) c = (char)tolower((unsigned char)c);
    for (const char* y : {"yes", "on", "true", "y", "t", "1"})
        if (t == y) return true;
    for (const char* n : {"no", "off", "false", "n", "f", "0", "none"})
        if (t == n) return false;
    fail("expected a switch for " + what + ", got '" + s + "'");
}
inline std::string expect(const std::string& v, std::initializer_list<const char*> allowed, const std::string& what) {
    std::string all;
    for (const char* a : allowed) {
        if (v == a) return v;
        all += (all.empty() ? "" : " | ") + std::string(a);
    }
    fail(what + ": '" + v + "' is not supported by this solver (supported: " + all + ")");
}

// ----------------------------------------------------------------------------------------
// tokenizer / recursive-descent parser (ascii and binary FoamFile streams)
// ----------------------------------------------------------------------------------------
struct Parser {
    const std::string& b;
    std::string name;
    size_t i = 0;
    bool binary = false;
    int labelBytes = 4, scalarBytes = 8;
    Dict header;

    Parser(const std::string& buf, const std::string& nm) : b(buf), name(nm) {}
    static bool ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }
    static bool punct(char c) { return c == '{' || c == '}' || c == '(' || c == ')' || c == ';'; }

    void skip() {
        size_t n = b.size();
        while (i < n) {
            char c = b[i];
            if (ws(c)) i++;
            else if (c == '/' && i + 1 < n && b[i + 1] == '/') {
                size_t j = b.find('\n', i);
                i = j == std::string::npos ? n : j + 1;
            } else if (c == '/' && i + 1 < n && b[i + 1] == '*') {
                size_t j = b.find("*/", i + 2);
                if (j == std::string::npos) fail(name + ": unterminated comment");
                i = j + 2;
            } else break;
        }
    }
    // one token without consuming it: "" at the end of the file
    std::string peek(size_t* endp = nullptr) {
        skip();
        size_t n = b.size();
        if (i >= n) return "";
        char c = b[i];
        if (punct(c)) {
            if (endp) *endp = i + 1;
            return std::string(1, c);
        }
        size_t j = i;
        if (c == '"') {
            j = i + 1;
            while (j < n && b[j] != '"') j += b[j] == '\\' ? 2 : 1;
            j = std::min(j + 1, n);
        } else {
            int depth = 0;
            bool digits = true;
            for (; j < n; j++) {
                char q = b[j];
                if (ws(q)) break;
                if (q == '(') {  // inside a word such as div(phi,alpha); `3(` opens a counted list
                    if (digits) break;
                    depth++;
                } else if (q == ')') {
                    if (depth == 0) break;
                    depth--;
                } else if ((q == '{' || q == '}' || q == ';') && depth == 0) break;
                if (q < '0' || q > '9') digits = false;
            }
        }
        if (endp) *endp = j;
        return b.substr(i, j - i);
    }
    std::string next() {
        size_t e = i;
        std::string t = peek(&e);
        if (!t.empty()) i = e;
        return t;
    }
    void expectTok(const char* t) {
        std::string g = next();
        if (g != t) fail(name + ": expected '" + t + "' got '" + g + "' near byte " + std::to_string(i));
    }
    static bool isInt(const std::string& t) { return !t.empty() && std::all_of(t.begin(), t.end(), [](char c) { return c >= '0' && c <= '9'; }); }

    void parseHeader() {
        if (peek() != "FoamFile") return;
        next();
        expectTok("{");
        header = parseDictBody(false);
        binary = wordOr(header, "format", "ascii") == "binary";
        std::string arch = wordOr(header, "arch", "");
        size_t k = arch.find("label=");
        if (k != std::string::npos) labelBytes = atoi(arch.c_str() + k + 6) / 8;
        k = arch.find("scalar=");
        if (k != std::string::npos) scalarBytes = atoi(arch.c_str() + k + 7) / 8;
        if ((labelBytes != 4 && labelBytes != 8) || (scalarBytes != 4 && scalarBytes != 8)) fail(name + ": unsupported arch " + arch);
    }
    Dict parseDictBody(bool top) {
        Dict d;
        for (;;) {
            std::string t = peek();
            if (t.empty()) {
                if (top) return d;
                fail(name + ": unexpected end of file in dictionary");
            }
            if (t == "}") {
                if (top) fail(name + ": unbalanced '}'");
                next();
                return d;
            }
            std::string key = next();
            if (key == "(" || key == ")" || key == "{" || key == ";") fail(name + ": unexpected '" + key + "' where a keyword was expected");
            if (key[0] == '#') fail(name + ": directive '" + key + "' is not supported");
            if (peek() == "{") {
                next();
                Node v;
                v.kind = Node::DICT;
                v.d = parseDictBody(false);
                d.emplace_back(key, Stream{std::move(v)});
                continue;
            }
            d.emplace_back(key, parseStream());
        }
    }
    Stream parseStream() {
        Stream out;
        for (;;) {
            std::string t = peek();
            if (t.empty()) fail(name + ": missing ';'");
            if (t == ";") {
                next();
                return out;
            }
            if (t == "{") {
                next();
                Node v;
                v.kind = Node::DICT;
                v.d = parseDictBody(false);
                out.push_back(std::move(v));
                continue;
            }
            out.push_back(parseValue(out));
        }
    }
    // the element type of `List<T>`: components, or 0 when T is not numeric
    static int listType(const std::string& t, bool& label) {
        label = t == "List<label>";
        if (label || t == "List<scalar>") return 1;
        if (t == "List<vector>") return 3;
        if (t == "List<symmTensor>") return 6;
        if (t == "List<tensor>") return 9;
        return 0;
    }
    // n x nc numbers; the stream stands on the opening parenthesis
    Node numList(long n, int nc, bool label) {
        Node v;
        v.kind = Node::NUM;
        v.n = n;
        v.nc = nc;
        // every value takes at least two bytes of the file (ascii) or its width (binary): refuse a count the
        // file cannot hold before allocating for it
        if (n < 0 || (double)n * nc * 2.0 > (double)b.size() + 2.0) fail(name + ": list of " + std::to_string(n) + " entries does not fit in the file");
        v.a.resize((size_t)n * nc);
        skip();
        if (i >= b.size() || b[i] != '(') fail(name + ": expected '(' before the data of a list of " + std::to_string(n));
        i++;
        if (binary) {
            size_t width = label ? labelBytes : (size_t)scalarBytes * nc, end = i + (size_t)n * width;
            if (end >= b.size() || b[end] != ')') fail(name + ": binary list of " + std::to_string(n) + " x " + std::to_string(width) + " bytes is truncated");
            const char* p = b.data() + i;
            size_t cnt = (size_t)n * nc;
            if (label) {
                if (labelBytes == 4) for (size_t k = 0; k < cnt; k++) { int32_t x; memcpy(&x, p + 4 * k, 4); v.a[k] = x; }
                else for (size_t k = 0; k < cnt; k++) { int64_t x; memcpy(&x, p + 8 * k, 8); v.a[k] = (double)x; }
            } else if (scalarBytes == 8) memcpy(v.a.data(), p, cnt * 8);
            else for (size_t k = 0; k < cnt; k++) { float x; memcpy(&x, p + 4 * k, 4); v.a[k] = x; }
            i = end + 1;
            return v;
        }
        size_t cnt = (size_t)n * nc, k = 0;
        const char* base = b.c_str();
        while (k < cnt) {
            while (i < b.size() && (ws(b[i]) || b[i] == '(' || b[i] == ')')) {
                // a ')' that closes the list before all numbers were seen
                if (b[i] == ')' && nc == 1) fail(name + ": list size mismatch: header says " + std::to_string(n) + ", found " + std::to_string(k) + " numbers");
                i++;
            }
            char* e = nullptr;
            double x = strtod(base + i, &e);
            if (e == base + i) fail(name + ": list size mismatch: header says " + std::to_string(n) + " x " + std::to_string(nc) + ", found " + std::to_string(k) + " numbers");
            v.a[k++] = x;
            i = e - base;
        }
        // the parentheses still open: the last element's (nc > 1) and the list's
        int closing = nc > 1 && n > 0 ? 2 : 1;
        while (closing) {
            skip();
            if (i >= b.size() || b[i] != ')') fail(name + ": list size mismatch: header says " + std::to_string(n) + " x " + std::to_string(nc) + ", found more");
            i++;
            closing--;
        }
        return v;
    }
    Node parseValue(const Stream& before) {
        std::string t = next();
        if (t == "(") return parseList();
        if (t.empty() || t == ")" || t == "}" || t == ";") fail(name + ": unexpected '" + t + "'");
        if (isInt(t)) {
            skip();
            char nxt = i < b.size() ? b[i] : 0;
            if (nxt == '(') {
                bool label = false;
                int nc = !before.empty() && before.back().kind == Node::WORD ? listType(before.back().w, label) : 0;
                if (nc) return numList(atol(t.c_str()), nc, label);
                // anything else (`inGroups List<word> 1(wall)` of a polyMesh/boundary file) is text in both formats
                i++;
                return parseList();
            }
            if (nxt == '{') fail(name + ": the N{value} list form is not supported");
        }
        Node v;
        v.w = t;
        return v;
    }
    Node parseList() {
        Node out;
        out.kind = Node::LIST;
        for (;;) {
            std::string t = peek();
            if (t.empty()) fail(name + ": unterminated list");
            if (t == ")") {
                next();
                return out;
            }
            if (t == "{") {  // list of named dictionaries: `name { ... }`
                next();
                Node v;
                v.kind = Node::DICT;
                v.d = parseDictBody(false);
                if (!out.l.empty() && out.l.back().kind == Node::WORD) {
                    v.w = out.l.back().w;
                    out.l.pop_back();
                }
                out.l.push_back(std::move(v));
                continue;
            }
            out.l.push_back(parseValue(out.l));
        }
    }
};

struct File {
    std::string path, buf;
    Parser p;
    explicit File(const std::string& pth) : path(pth), buf(slurp(pth)), p(buf, pth) { p.parseHeader(); }
};
inline Dict readDict(const std::string& path) {
    File f(path);
    return f.p.parseDictBody(true);
}

// ----------------------------------------------------------------------------------------
// polyMesh
// ----------------------------------------------------------------------------------------
struct Patch {
    std::string name, type;
    int nFaces = 0, startFace = 0, neighbProc = -1;
};
struct Mesh {
    std::vector<double> points;
    std::vector<int> fOff, fLab, owner, neighbour;
    std::vector<Patch> patches;
    std::map<std::string, long> cellZoneSizes;
    int nCells = 0;
    int nInternal() const { return (int)neighbour.size(); }
    int nFaces() const { return (int)owner.size(); }
};
inline Node readCounted(File& f, int nc, bool label) {
    std::string t = f.p.next();
    if (!Parser::isInt(t)) fail(f.path + ": expected 'N (' ");
    return f.p.numList(atol(t.c_str()), nc, label);
}
inline std::vector<int> toInts(const Node& v, const std::string& what) {
    std::vector<int> o(v.a.size());
    for (size_t k = 0; k < o.size(); k++) {
        if (v.a[k] < -1 || v.a[k] > 2147483647.0) fail(what + ": label out of range");
        o[k] = (int)v.a[k];
    }
    return o;
}
inline void readFaces(const std::string& path, std::vector<int>& off, std::vector<int>& lab) {
    File f(path);
    if (wordOr(f.p.header, "class", "") == "faceCompactList") {
        off = toInts(readCounted(f, 1, true), path);
        lab = toInts(readCounted(f, 1, true), path);
        return;
    }
    if (f.p.binary) fail(path + ": a binary faceList is not supported (faceCompactList is)");
    std::string t = f.p.next();  // ascii faceList: N ( 3(a b c) 4(a b c d) ... )
    if (!Parser::isInt(t)) fail(path + ": expected 'N (' ");
    long n = atol(t.c_str());
    f.p.expectTok("(");
    off.assign(1, 0);
    lab.clear();
    for (long k = 0; k < n; k++) {
        std::string s = f.p.next();
        if (!Parser::isInt(s)) fail(path + ": face " + std::to_string(k) + ": expected a point count, got '" + s + "'");
        f.p.expectTok("(");
        for (long q = 0, m = atol(s.c_str()); q < m; q++) {
            std::string l = f.p.next();
            if (!Parser::isInt(l)) fail(path + ": face " + std::to_string(k) + ": expected a point label, got '" + l + "'");
            lab.push_back(atoi(l.c_str()));
        }
        f.p.expectTok(")");
        off.push_back((int)lab.size());
    }
    f.p.expectTok(")");
}
inline Mesh readPolyMesh(const std::string& caseDir) {
    std::string d = caseDir + "/constant/polyMesh";
    if (!isDir(d)) fail(d + ": no polyMesh (run gmshToFoam or the repo's mesh generator first)");
    Mesh m;
    {
        File f(d + "/points");
        m.points = readCounted(f, 3, false).a;
    }
    readFaces(d + "/faces", m.fOff, m.fLab);
    {
        File f(d + "/owner");
        m.owner = toInts(readCounted(f, 1, true), f.path);
    }
    {
        File f(d + "/neighbour");
        m.neighbour = toInts(readCounted(f, 1, true), f.path);
    }
    {
        File f(d + "/boundary");
        Stream none;
        Node v = f.p.parseValue(none);
        if (v.kind != Node::LIST) fail(f.path + ": expected a list of patches");
        for (auto& it : v.l) {
            if (it.kind != Node::DICT) continue;
            Patch q;
            q.name = it.w;
            q.type = wordOr(it.d, "type", "patch");
            q.nFaces = integerOr(it.d, "nFaces", f.path + ":" + q.name, -1);
            q.startFace = integerOr(it.d, "startFace", f.path + ":" + q.name, -1);
            if (q.nFaces < 0 || q.startFace < 0) fail(f.path + ":" + q.name + ": nFaces / startFace missing");
            q.neighbProc = integerOr(it.d, "neighbProcNo", f.path + ":" + q.name, -1);
            m.patches.push_back(q);
        }
    }
    if (exists(d + "/cellZones")) {
        File f(d + "/cellZones");
        Stream none;
        Node v = f.p.parseValue(none);
        for (auto& it : v.l)
            if (it.kind == Node::DICT) {
                const Stream* cl = find(it.d, "cellLabels");
                if (cl && !cl->empty() && cl->back().kind == Node::NUM) m.cellZoneSizes[it.w] = cl->back().n;
                else if (cl && !cl->empty() && cl->back().kind == Node::LIST) m.cellZoneSizes[it.w] = (long)cl->back().l.size();
            }
    }
    // consistency (PolyMesh.check)
    int nF = m.nFaces(), nI = m.nInternal(), nP = (int)(m.points.size() / 3);
    if ((int)m.fOff.size() != nF + 1) fail(d + ": faces and owner disagree on the number of faces");
    if (nI > nF) fail(d + ": more neighbours than faces");
    // the array ABI carries no length for face_labels: the offsets are checked against it here
    if (m.fOff[0] != 0 || m.fOff[nF] != (int)m.fLab.size()) fail(d + "/faces: the offsets do not span the " + std::to_string(m.fLab.size()) + " point labels");
    for (int f = 0; f < nF; f++)
        if (m.fOff[f + 1] < m.fOff[f]) fail(d + "/faces: offsets of face " + std::to_string(f) + " decrease");
    for (int l : m.fLab)
        if (l < 0 || l >= nP) fail(d + "/faces: point label out of range");
    int mx = -1;
    for (int c : m.owner) {
        if (c < 0) fail(d + "/owner: negative cell label");
        mx = std::max(mx, c);
    }
    for (int c : m.neighbour) mx = std::max(mx, c);
    m.nCells = mx + 1;
    int at = nI;
    for (auto& q : m.patches) {
        if (q.startFace != at) fail(d + "/boundary: patch '" + q.name + "' does not start where the previous one ends");
        at += q.nFaces;
    }
    if (at != nF) fail(d + "/boundary: patches cover " + std::to_string(at - nI) + " of " + std::to_string(nF - nI) + " boundary faces");
    return m;
}

// ----------------------------------------------------------------------------------------
// fields
// ----------------------------------------------------------------------------------------
struct Value {  // `uniform X` or `nonuniform List<T> N (...)`
    bool present = false, uniform = true;
    double u[9] = {0};
    std::vector<double> a;
    // n values of nc components, whichever form the file used
    std::vector<double> expand(long n, int nc, const std::string& what) const {
        std::vector<double> o((size_t)n * nc);
        if (uniform) {
            for (long k = 0; k < n; k++)
                for (int c = 0; c < nc; c++) o[(size_t)k * nc + c] = u[c];
        } else {
            if ((long)a.size() != n * nc) fail(what + ": holds " + std::to_string(a.size() / nc) + " values, the mesh needs " + std::to_string(n));
            o = a;
        }
        return o;
    }
};
struct BoundaryEntry {
    std::string patch;
    std::vector<std::pair<std::string, std::string>> entries;  // everything but `value`, as text
    Value value;
    const std::string* get(const char* k) const {
        for (auto& e : entries)
            if (e.first == k) return &e.second;
        return nullptr;
    }
};
struct Field {
    std::string cls, dimensions;
    int nc = 1;
    Value internal;
    std::vector<BoundaryEntry> boundary;
    const BoundaryEntry* patch(const std::string& nm) const {
        for (auto& b : boundary)
            if (b.patch == nm) return &b;
        return nullptr;
    }
};
inline Value fieldValue(Stream& s, int nc, const std::string& what) {  // takes the numbers out of the parse tree
    Value v;
    v.present = true;
    if (s.size() >= 2 && s[0].kind == Node::WORD && s[0].w == "uniform") {
        if (nc == 1 && s[1].kind == Node::WORD) v.u[0] = toNumber(s[1].w, what);
        else if (nc > 1 && s[1].kind == Node::LIST && (int)s[1].l.size() == nc)
            for (int c = 0; c < nc; c++) v.u[c] = toNumber(s[1].l[c].w, what);
        else fail(what + ": malformed uniform value '" + join(s) + "'");
        return v;
    }
    if (s.size() == 3 && s[0].kind == Node::WORD && s[0].w == "nonuniform" && s[2].kind == Node::NUM && s[2].nc == nc) {
        v.uniform = false;
        v.a = std::move(s[2].a);
        return v;
    }
    fail(what + ": unsupported field value '" + join(s).substr(0, 60) + "'");
}
inline Field readField(const std::string& path) {
    File f(path);
    Field fld;
    fld.cls = wordOr(f.p.header, "class", "");
    if (fld.cls == "volScalarField" || fld.cls == "surfaceScalarField") fld.nc = 1;
    else if (fld.cls == "volVectorField" || fld.cls == "surfaceVectorField") fld.nc = 3;
    else fail(path + ": unsupported field class '" + fld.cls + "'");
    Dict d = f.p.parseDictBody(true);
    fld.dimensions = wordOr(d, "dimensions", "");
    fld.internal = fieldValue(const_cast<Stream&>(lookup(d, "internalField", path)), fld.nc, path + ":internalField");
    if (Stream* bf = const_cast<Stream*>(find(d, "boundaryField"))) {
        if (bf->size() != 1 || (*bf)[0].kind != Node::DICT) fail(path + ": boundaryField is not a dictionary");
        for (auto& kv : (*bf)[0].d) {
            BoundaryEntry e;
            e.patch = kv.first;
            if (kv.second.size() != 1 || kv.second[0].kind != Node::DICT) fail(path + ":" + kv.first + ": not a dictionary");
            for (auto& q : kv.second[0].d) {
                if (q.first == "value") e.value = fieldValue(q.second, fld.nc, path + ":" + kv.first + ".value");
                else e.entries.emplace_back(q.first, join(q.second));
            }
            fld.boundary.push_back(std::move(e));
        }
    }
    return fld;
}

// ----------------------------------------------------------------------------------------
// writers (what foamRun leaves in a time directory; controlDict:35-37 writeFormat / writePrecision)
// ----------------------------------------------------------------------------------------
inline std::string fileHeader(const char* cls, const std::string& obj, const std::string& location, bool binary) {
    std::string s =
        "/*--------------------------------*- C++ -*----------------------------------*\\\n"
        "  =========                 |\n"
        "  \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox\n"
        "   \\\\    /   O peration     | Website:  https://openfoam.org\n"
        "    \\\\  /    A nd           | Version:  13\n"
        "     \\\\/     M anipulation  |\n"
        "\\*---------------------------------------------------------------------------*/\n"
        "FoamFile\n{\n";
    s += std::string("    format      ") + (binary ? "binary" : "ascii") + ";\n";
    s += std::string("    class       ") + cls + ";\n";
    if (binary) s += "    arch        \"LSB;label=32;scalar=64\";\n";
    if (!location.empty()) s += "    location    \"" + location + "\";\n";
    s += "    object      " + obj + ";\n}\n";
    s += "// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n";
    return s;
}
static const char* const FILE_END = "\n// ************************************************************************* //\n";

struct Out {
    FILE* f;
    std::string path;
    explicit Out(const std::string& p) : f(fopen(p.c_str(), "wb")), path(p) {
        if (!f) fail(p + ": cannot open for writing");
    }
    ~Out() {
        if (f) fclose(f);
    }
    void str(const std::string& s) {
        if (fwrite(s.data(), 1, s.size(), f) != s.size()) fail(path + ": write failed");
    }
    void close() {
        FILE* g = f;
        f = nullptr;
        if (fclose(g) != 0) fail(path + ": write failed");
    }
};
inline std::string fmtNum(double x, int prec) {
    char b[64];
    snprintf(b, sizeof b, "%.*g", prec, x);
    return b;
}
// `nonuniform List<T> N ( ... )` of n values with nc components
inline void writeList(Out& o, const double* a, long n, int nc, bool binary, int prec) {
    o.str(std::string("nonuniform List<") + (nc == 1 ? "scalar" : "vector") + "> \n" + std::to_string(n) + "\n(");
    if (binary) {
        if (n && fwrite(a, sizeof(double), (size_t)n * nc, o.f) != (size_t)n * nc) fail(o.path + ": write failed");
        o.str(")");
        return;
    }
    std::string s = "\n";
    for (long k = 0; k < n; k++) {
        if (k) s += "\n";
        if (nc == 1) s += fmtNum(a[k], prec);
        else {
            s += "(";
            for (int c = 0; c < nc; c++) s += (c ? " " : "") + fmtNum(a[(size_t)k * nc + c], prec);
            s += ")";
        }
        if (s.size() > (1 << 16)) {
            o.str(s);
            s.clear();
        }
    }
    s += "\n)";
    o.str(s);
}
struct PatchOut {
    std::string name;
    std::vector<std::pair<std::string, std::string>> entries;
    const double* value = nullptr;
    long n = 0;
};
inline void writeField(const std::string& path, const char* cls, const std::string& obj, const std::string& location, const std::string& dims, const double* internal, long n, int nc,
                       const std::vector<PatchOut>& patches, bool binary, int prec) {
    Out o(path);
    o.str(fileHeader(cls, obj, location, binary));
    o.str("dimensions      " + dims + ";\n\ninternalField   ");
    writeList(o, internal, n, nc, binary, prec);
    o.str(";\n\nboundaryField\n{\n");
    for (auto& p : patches) {
        o.str("    " + p.name + "\n    {\n");
        for (auto& e : p.entries) {
            char b[64];
            snprintf(b, sizeof b, "        %-15s ", e.first.c_str());
            o.str(std::string(b) + e.second + ";\n");
        }
        if (p.value || p.n == 0) {
            o.str("        value           ");
            writeList(o, p.value, p.n, nc, binary, prec);
            o.str(";\n");
        }
        o.str("    }\n");
    }
    o.str("}\n");
    o.str(FILE_END);
    o.close();
}
inline void writePoints(const std::string& path, const std::string& location, const double* pts, long n, bool binary) {
    Out o(path);
    o.str(fileHeader("vectorField", "points", location, binary));
    o.str("\n" + std::to_string(n) + "\n(");
    if (binary) {
        if (n && fwrite(pts, sizeof(double), (size_t)3 * n, o.f) != (size_t)3 * n) fail(path + ": write failed");
        o.str(")\n");
    } else {
        std::string s = "\n";
        for (long k = 0; k < n; k++) {
            s += "(" + fmtNum(pts[3 * k], 17) + " " + fmtNum(pts[3 * k + 1], 17) + " " + fmtNum(pts[3 * k + 2], 17) + ")" + (k + 1 < n ? "\n" : "");
            if (s.size() > (1 << 16)) {
                o.str(s);
                s.clear();
            }
        }
        o.str(s + "\n)\n");
    }
    o.str(FILE_END);
    o.close();
}

// Time::timeName with `timeFormat general`: ostream << setprecision(p)
inline std::string timeName(double t, int precision) { return fmtNum(t, precision); }

// numeric directories of a case, sorted by value
inline std::vector<std::pair<double, std::string>> timeDirs(const std::string& caseDir) {
    std::vector<std::pair<double, std::string>> out;
    DIR* d = opendir(caseDir.c_str());
    if (!d) fail(caseDir + ": cannot list the case directory");
    while (dirent* e = readdir(d)) {
        std::string nm = e->d_name;
        if (nm.empty() || !(isdigit((unsigned char)nm[0]) || nm[0] == '.' || nm[0] == '-' || nm[0] == '+')) continue;
        char* end = nullptr;
        double v = strtod(nm.c_str(), &end);
        if (end == nm.c_str() || *end || !isDir(caseDir + "/" + nm)) continue;
        out.emplace_back(v, nm);
    }
    closedir(d);
    std::sort(out.begin(), out.end());
    return out;
}
// newest COMPLETE time directory (case.latest_time): a directory the solver wrote counts once its
// uniform/time exists (written last); a hand-made start directory counts as soon as it holds alpha.water
inline std::pair<double, std::string> latestTime(const std::string& caseDir) {
    bool any = false;
    std::pair<double, std::string> best;
    for (auto& tn : timeDirs(caseDir)) {
        std::string d = caseDir + "/" + tn.second;
        if (!exists(d + "/alpha.water")) continue;
        if (tn.first > 0 && exists(d + "/phi") && !exists(d + "/uniform/time")) continue;
        best = tn;
        any = true;
    }
    if (!any) fail(caseDir + ": no time directory holds alpha.water");
    return best;
}

// constant/6DoF.dat (generate_motion.py:13-42): N ( (t (tx ty tz) (rx ry rz)) ... )
inline std::vector<double> readMotionTable(const std::string& path) {
    std::string raw = slurp(path), buf;
    for (size_t i = 0; i < raw.size();) {  // strip // comments
        if (raw[i] == '/' && i + 1 < raw.size() && raw[i + 1] == '/') {
            while (i < raw.size() && raw[i] != '\n') i++;
        } else buf += raw[i++];
    }
    const char* p = buf.c_str();
    char* e = nullptr;
    long n = strtol(p, &e, 10);
    if (e == p || n < 0) fail(path + ": expected 'N (' at the top of the motion table");
    p = e;
    while (*p && Parser::ws(*p)) p++;
    if (*p != '(') fail(path + ": expected 'N (' at the top of the motion table");
    std::vector<double> a;
    for (;;) {
        while (*p && (Parser::ws(*p) || *p == '(' || *p == ')')) p++;
        if (!*p) break;
        double v = strtod(p, &e);
        if (e == p) fail(path + ": unexpected text in the motion table");
        a.push_back(v);
        p = e;
    }
    if ((long)a.size() != 7 * n) fail(path + ": table says " + std::to_string(n) + " rows but holds " + std::to_string(a.size()) + " numbers (expected " + std::to_string(7 * n) + ")");
    for (long k = 1; k < n; k++)
        if (!(a[7 * k] > a[7 * (k - 1)])) fail(path + ": table times are not strictly increasing");
    return a;
}

// ----------------------------------------------------------------------------------------
// the dictionaries -> tpp_config_t   (case.read_config)
// ----------------------------------------------------------------------------------------
struct Config {
    tpp_config_t c;
    std::vector<double> motion;
    bool writeBinary = false;
    int writePrecision = 6, timePrecision = 6;
    bool startLatest = true;
    double startTimeEntry = 0;
    std::vector<double> probes;  // n x 3
    bool hasProbes = false;
};
inline int smootherCode(const std::string& sm, const std::string& where) {
    if (sm == "DIC") return 0;
    if (sm == "DICGaussSeidel") return 1;
    if (sm == "GaussSeidel") return 2;
    fail(where + ": smoother '" + sm + "' is not supported (DIC, DICGaussSeidel, GaussSeidel)");
}
inline void gamgOptions(const Dict& g, tpp_solver_t& sc, const std::string& where) {
    sc.smoother = smootherCode(word(g, "smoother", where), where);
    sc.n_vcycles = integerOr(g, "nVcycles", where, 2);
    sc.n_pre_sweeps = integerOr(g, "nPreSweeps", where, 0);
    sc.n_post_sweeps = integerOr(g, "nPostSweeps", where, 2);
    sc.n_finest_sweeps = integerOr(g, "nFinestSweeps", where, 2);
    sc.n_cells_coarsest = integerOr(g, "nCellsInCoarsestLevel", where, 10);
    sc.merge_levels = integerOr(g, "mergeLevels", where, 1);
    std::string agg = wordOr(g, "agglomerator", "faceAreaPair");
    if (agg != "faceAreaPair") fail(where + ": agglomerator '" + agg + "' is not supported (faceAreaPair)");
}
inline tpp_solver_t solverControl(const Dict& d, const std::string& name, const std::string& path) {
    tpp_solver_t sc;
    memset(&sc, 0, sizeof sc);
    sc.n_vcycles = 2, sc.n_post_sweeps = 2, sc.n_finest_sweeps = 2, sc.n_cells_coarsest = 10, sc.merge_levels = 1;
    std::string where = path + ":" + name, solver = word(d, "solver", where);
    sc.tolerance = number(d, "tolerance", where);
    sc.rel_tol = numberOr(d, "relTol", where, 0.0);
    sc.max_iter = integerOr(d, "maxIter", where, 1000);
    if (solver == "GAMG") {
        sc.type = 1;
        gamgOptions(d, sc, where);
    } else if (solver == "PCG") {
        sc.type = 0;
        const Stream& pre = lookup(d, "preconditioner", where);
        if (pre.size() == 1 && pre[0].kind == Node::DICT) {
            std::string kind = word(pre[0].d, "preconditioner", where + ".preconditioner");
            if (kind == "GAMG") {
                sc.precond = 1;
                gamgOptions(pre[0].d, sc, where);
            } else if (kind == "DIC") sc.precond = 0;
            else fail(where + ": preconditioner '" + kind + "' is not supported (GAMG, DIC)");
        } else if (join(pre) == "DIC") sc.precond = 0;
        else fail(where + ": preconditioner '" + join(pre) + "' is not supported (GAMG, DIC)");
    } else fail(where + ": solver '" + solver + "' is not supported (PCG, GAMG)");
    return sc;
}
inline std::string replaceAll(std::string s, const std::string& a, const std::string& b) {
    for (size_t k = 0; (k = s.find(a, k)) != std::string::npos; k += b.size()) s.replace(k, a.size(), b);
    return s;
}
inline void readConfig(const std::string& root, const Mesh& mesh, Config& cfg) {
    tpp_config_t& c = cfg.c;
    memset(&c, 0, sizeof c);
    std::string p = root + "/system/controlDict";
    Dict cd = readDict(p);
    expect(word(cd, "solver", p), {"incompressibleVoF"}, p + ":solver");
    c.end_time = number(cd, "endTime", p);
    c.delta_t = number(cd, "deltaT", p);
    expect(word(cd, "writeControl", p), {"adjustableRunTime"}, p + ":writeControl");
    c.write_interval = number(cd, "writeInterval", p);
    c.adjust_time_step = toSwitch(wordOr(cd, "adjustTimeStep", "no"), p + ":adjustTimeStep");
    c.max_co = numberOr(cd, "maxCo", p, 1.0);
    c.max_alpha_co = numberOr(cd, "maxAlphaCo", p, 1.0);
    c.max_delta_t = numberOr(cd, "maxDeltaT", p, 1e30);
    cfg.writeBinary = wordOr(cd, "writeFormat", "ascii") == "binary";
    cfg.writePrecision = integerOr(cd, "writePrecision", p, 6);
    cfg.timePrecision = integerOr(cd, "timePrecision", p, 6);
    expect(wordOr(cd, "timeFormat", "general"), {"general"}, p + ":timeFormat");
    cfg.startLatest = expect(wordOr(cd, "startFrom", "latestTime"), {"latestTime", "startTime"}, p + ":startFrom") == "latestTime";
    cfg.startTimeEntry = numberOr(cd, "startTime", p, 0.0);

    p = root + "/system/fvSchemes";
    Dict fs = readDict(p);
    expect(word(subDict(fs, "ddtSchemes", p), "default", p), {"Euler"}, p + ":ddtSchemes");
    expect(word(subDict(fs, "gradSchemes", p), "default", p), {"Gauss linear"}, p + ":gradSchemes");
    const Dict& div = subDict(fs, "divSchemes", p);
    expect(word(div, "div(rhoPhi,U)", p), {"Gauss vanLeerV"}, p + ":div(rhoPhi,U)");
    const Stream& da = lookup(div, "div(phi,alpha)", p);
    if (!(da.size() == 4 && da[0].w == "Gauss" && da[1].w == "interfaceCompression" && da[2].w == "vanLeer" && da[3].kind == Node::WORD))
        fail(p + ":div(phi,alpha): only 'Gauss interfaceCompression vanLeer <cAlpha>' is supported, got '" + join(da) + "'");
    c.c_alpha = toNumber(da[3].w, p + ":div(phi,alpha)");
    expect(word(div, "div(((rho*nuEff)*dev2(T(grad(U)))))", p), {"Gauss linear"}, p + ":div(((rho*nuEff)*dev2(T(grad(U)))))");
    expect(word(subDict(fs, "laplacianSchemes", p), "default", p), {"Gauss linear corrected"}, p + ":laplacianSchemes");
    expect(word(subDict(fs, "interpolationSchemes", p), "default", p), {"linear"}, p + ":interpolationSchemes");
    expect(word(subDict(fs, "snGradSchemes", p), "default", p), {"corrected"}, p + ":snGradSchemes");

    p = root + "/system/fvSolution";
    Dict fv = readDict(p);
    const Dict& sol = subDict(fv, "solvers", p);
    const Dict& a = subDict(sol, "alpha.water", p);
    c.n_alpha_subcycles = integerOr(a, "nSubCycles", p, integerOr(a, "nAlphaSubCycles", p, 1));
    c.n_alpha_corr = integerOr(a, "nCorrectors", p, integerOr(a, "nAlphaCorr", p, 1));
    c.n_limiter_iter = integerOr(a, "nLimiterIter", p, 3);
    if (toSwitch(wordOr(a, "MULESCorr", "no"), p + ":MULESCorr")) fail(p + ":alpha.water: MULESCorr yes (semi-implicit MULES) is not supported");
    c.p_rgh = solverControl(subDict(sol, "p_rgh", p), "p_rgh", p);
    c.p_rgh_final = solverControl(subDict(sol, "p_rghFinal", p), "p_rghFinal", p);
    const Dict& pim = subDict(fv, "PIMPLE", p);
    if (toSwitch(wordOr(pim, "momentumPredictor", "yes"), p + ":momentumPredictor")) fail(p + ":PIMPLE: momentumPredictor yes is not supported (the reference runs with 'no')");
    if (integerOr(pim, "nOuterCorrectors", p, 1) != 1) fail(p + ":PIMPLE: nOuterCorrectors != 1 is not supported");
    if (toSwitch(wordOr(pim, "correctPhi", "yes"), p + ":correctPhi")) fail(p + ":PIMPLE: correctPhi yes is not supported (the reference runs with 'no')");
    c.n_correctors = integerOr(pim, "nCorrectors", p, 1);
    c.n_non_orth = integerOr(pim, "nNonOrthogonalCorrectors", p, 0);
    if (const Stream* rp = find(pim, "pRefPoint")) {
        vector3(*rp, p + ":pRefPoint", c.p_ref_point);
        c.p_ref_value = number(pim, "pRefValue", p);
    }

    p = root + "/constant/g";
    vector3(lookup(readDict(p), "value", p), p + ":value", c.g);
    p = root + "/constant/momentumTransport";
    expect(word(readDict(p), "simulationType", p), {"laminar"}, p + ":simulationType");
    p = root + "/constant/phaseProperties";
    Dict pp = readDict(p);
    const Stream& ph = lookup(pp, "phases", p);
    if (join(ph) != "(water air)") fail(p + ":phases: expected (water air), got " + join(ph));
    c.sigma = number(pp, "sigma", p);
    if (!(c.sigma >= 0.0 && std::isfinite(c.sigma))) fail(p + ":sigma: expected a non-negative surface tension coefficient");
    for (int k = 0; k < 2; k++) {
        p = root + "/constant/physicalProperties." + (k ? "air" : "water");
        Dict d = readDict(p);
        expect(word(d, "viscosityModel", p), {"constant"}, p + ":viscosityModel");
        (k ? c.rho2 : c.rho1) = number(d, "rho", p);
        (k ? c.nu2 : c.nu1) = number(d, "nu", p);
    }

    p = root + "/constant/dynamicMeshDict";
    if (exists(p)) {
        Dict dm = readDict(p);
        const Dict& mv = subDict(dm, "mover", p);
        expect(word(mv, "motionSolver", p), {"solidBody"}, p + ":motionSolver");
        expect(word(mv, "solidBodyMotionFunction", p), {"sixDoFMotion"}, p + ":solidBodyMotionFunction");
        vector3(lookup(mv, "CofG", p), p + ":CofG", c.cofg);
        std::string file;
        for (int k = 0; k < 2; k++) {
            const char* key = k ? "rotation" : "translation";
            const Dict& e = subDict(mv, key, p);
            expect(word(e, "type", p), {"table"}, p + ":" + key + ".type");
            if (word(e, "columns", p) != (k ? "(0 2)" : "(0 1)")) fail(p + ":" + key + ".columns: expected " + (k ? "(0 2)" : "(0 1)"));
            std::string fn = word(e, "file", p);
            if (fn.size() >= 2 && fn[0] == '"') fn = fn.substr(1, fn.size() - 2);
            fn = replaceAll(fn, "$FOAM_CASE", root);
            if (k && fn != file) fail(p + ": translation and rotation must read the same table file");
            file = fn;
        }
        cfg.motion = readMotionTable(file);
        c.n_motion = (int)(cfg.motion.size() / 7);
        std::string zone = word(mv, "cellZone", p);
        if (!mesh.cellZoneSizes.empty()) {
            auto it = mesh.cellZoneSizes.find(zone);
            if (it == mesh.cellZoneSizes.end()) fail(p + ":cellZone '" + zone + "' is not in constant/polyMesh/cellZones");
            if (it->second != mesh.nCells) fail(p + ":cellZone '" + zone + "' does not cover the whole mesh; partial solid-body zones are not supported");
        }
    }

    p = root + "/system/functions";
    if (exists(p)) {
        Dict fo = readDict(p);
        for (auto& kv : fo) {
            if (kv.second.size() != 1 || kv.second[0].kind != Node::DICT) continue;
            const Dict& o = kv.second[0].d;
            std::string typ = word(o, "type", p);
            if (typ != "probes") fail(p + ":" + kv.first + ": function object type '" + typ + "' is not supported (probes)");
            const Stream& pl = lookup(o, "probeLocations", p);
            if (pl.empty() || pl.back().kind != Node::LIST) fail(p + ":" + kv.first + ": probeLocations is not a list");
            cfg.probes.clear();
            for (auto& pt : pl.back().l) {
                if (pt.kind != Node::LIST || pt.l.size() != 3) fail(p + ":" + kv.first + ": probeLocations entries must be (x y z)");
                for (int k = 0; k < 3; k++) cfg.probes.push_back(toNumber(pt.l[k].w, p + ":probeLocations"));
            }
            cfg.hasProbes = true;
            std::string flds = word(o, "fields", p);
            // the solver samples p (system/functions:28-31); any other selection would be written under the wrong name
            if (flds != "(p)") fail(p + ":" + kv.first + ": probes 'fields " + flds + "' is not supported: exactly 'fields (p)'");
        }
    }
}

// ----------------------------------------------------------------------------------------
// a case in memory
// ----------------------------------------------------------------------------------------
struct Case {
    std::string root, dir;  // dictionaries under root; mesh, fields and time directories under dir (root or root/processorN)
    Mesh mesh;
    Config cfg;
    double startValue = 0;
    std::string startName;
    Field alpha, U, p_rgh, phi, Uf;
    bool hasFlux = false, hasRestartDt = false;
    double restartDt = 0;
    std::vector<int> bcU, bcA, bcP, pStart, pSize, neighb;
    std::vector<double> inletAlpha, p0;
    // run state of tpp_run_case
    FILE* probesFile = nullptr;
    std::vector<int> probeCells;
    bool started = false, interfaceStarted = false;
    ~Case() {
        if (probesFile) fclose(probesFile);
    }
};
inline int bcCode(const std::string& fld, std::string t, const std::string& where) {
    if (fld == "U") {
        if (t == "noSlip") t = "movingWallVelocity";  // identical on a wall that moves with the mesh
        if (t == "movingWallVelocity") return TPP_U_MOVING_WALL;
        if (t == "pressureInletOutletVelocity") return TPP_U_PRESSURE_INLET_OUTLET;
        fail(where + ": boundary condition '" + t + "' is not supported (movingWallVelocity, noSlip, pressureInletOutletVelocity)");
    }
    if (fld == "alpha.water") {
        if (t == "zeroGradient") return TPP_A_ZERO_GRADIENT;
        if (t == "inletOutlet") return TPP_A_INLET_OUTLET;
        fail(where + ": boundary condition '" + t + "' is not supported (zeroGradient, inletOutlet)");
    }
    if (t == "fixedFluxPressure") return TPP_P_FIXED_FLUX;
    if (t == "totalPressure") return TPP_P_TOTAL_PRESSURE;
    fail(where + ": boundary condition '" + t + "' is not supported (fixedFluxPressure, totalPressure)");
}
inline double uniformEntry(const BoundaryEntry& e, const char* key, const std::string& where) {
    const std::string* s = e.get(key);
    if (!s) return 0.0;
    size_t k = s->rfind(' ');
    return toNumber(k == std::string::npos ? *s : s->substr(k + 1), where + ":" + key);
}
inline void load(const std::string& root, int processor, Case& cs) {
    cs.root = root;
    cs.dir = root;
    if (!isDir(root)) fail(root + ": not a case directory");
    if (processor >= 0) {
        cs.dir = root + "/processor" + std::to_string(processor);
        if (!isDir(cs.dir)) fail(cs.dir + ": not found (run decomposePar first)");
    }
    cs.mesh = readPolyMesh(cs.dir);
    readConfig(root, cs.mesh, cs.cfg);
    if (cs.cfg.startLatest) {
        auto lt = latestTime(cs.dir);
        cs.startValue = lt.first;
        cs.startName = lt.second;
    } else {
        cs.startValue = cs.cfg.startTimeEntry;
        cs.startName = timeName(cs.startValue, cs.cfg.timePrecision);
    }
    cs.cfg.c.start_time = cs.startValue;
    std::string tdir = cs.dir + "/" + cs.startName;
    cs.alpha = readField(tdir + "/alpha.water");
    cs.U = readField(tdir + "/U");
    cs.p_rgh = readField(tdir + "/p_rgh");
    if (cs.alpha.nc != 1 || cs.U.nc != 3 || cs.p_rgh.nc != 1) fail(tdir + ": alpha.water / U / p_rgh have the wrong field class");
    if (exists(tdir + "/phi") && exists(tdir + "/Uf")) {
        cs.phi = readField(tdir + "/phi");
        cs.Uf = readField(tdir + "/Uf");
        cs.hasFlux = true;
    }
    if (exists(tdir + "/uniform/time")) {
        std::string up = tdir + "/uniform/time";
        cs.restartDt = number(readDict(up), "deltaT", up);
        cs.hasRestartDt = true;
    }
    for (auto& q : cs.mesh.patches) {
        cs.pStart.push_back(q.startFace);
        cs.pSize.push_back(q.nFaces);
        cs.neighb.push_back(q.neighbProc);
        if (q.type == "processor") {
            cs.bcU.push_back(-1), cs.bcA.push_back(-1), cs.bcP.push_back(-1);
            cs.inletAlpha.push_back(0.0), cs.p0.push_back(0.0);
            continue;
        }
        const Field* flds[3] = {&cs.U, &cs.alpha, &cs.p_rgh};
        const char* names[3] = {"U", "alpha.water", "p_rgh"};
        std::vector<int>* outs[3] = {&cs.bcU, &cs.bcA, &cs.bcP};
        for (int k = 0; k < 3; k++) {
            std::string where = cs.startName + "/" + names[k] + ":" + q.name;
            const BoundaryEntry* e = flds[k]->patch(q.name);
            if (!e) fail("keyword '" + q.name + "' is missing in " + cs.startName + "/" + names[k] + ":boundaryField");
            const std::string* t = e->get("type");
            if (!t) fail("keyword 'type' is missing in " + where);
            outs[k]->push_back(bcCode(names[k], *t, where));
        }
        cs.inletAlpha.push_back(uniformEntry(*cs.alpha.patch(q.name), "inletValue", tdir + "/alpha.water:" + q.name));
        cs.p0.push_back(uniformEntry(*cs.p_rgh.patch(q.name), "p0", tdir + "/p_rgh:" + q.name));
    }
    cs.cfg.c.motion = cs.cfg.motion.empty() ? nullptr : cs.cfg.motion.data();
}
inline tpp_mesh_t meshView(const Case& cs) {
    tpp_mesh_t m;
    memset(&m, 0, sizeof m);
    m.n_points = (int)(cs.mesh.points.size() / 3);
    m.n_faces = cs.mesh.nFaces();
    m.n_internal = cs.mesh.nInternal();
    m.n_cells = cs.mesh.nCells;
    m.n_patches = (int)cs.mesh.patches.size();
    m.points = cs.mesh.points.data();
    m.face_offsets = cs.mesh.fOff.data();
    m.face_labels = cs.mesh.fLab.data();
    m.owner = cs.mesh.owner.data();
    m.neighbour = cs.mesh.neighbour.data();
    m.patch_start = cs.pStart.data();
    m.patch_size = cs.pSize.data();
    m.patch_bc_u = cs.bcU.data();
    m.patch_bc_alpha = cs.bcA.data();
    m.patch_bc_p = cs.bcP.data();
    m.patch_inlet_alpha = cs.inletAlpha.data();
    m.patch_p0 = cs.p0.data();
    m.patch_neighb_proc = cs.neighb.data();
    return m;
}

}  // namespace caseio
