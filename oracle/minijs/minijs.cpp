// minijs.cpp -- see minijs.h. TEST INFRASTRUCTURE ONLY (never part of the product path).
#include "minijs.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace mjs {

// =====================================================================================================
// Values
// =====================================================================================================
typedef uint32_t Atom;
typedef std::u16string U16;

struct Obj;
struct Str {
    uint32_t rc = 0;
    U16 s;
};

enum Tag : uint8_t { T_UNDEF, T_NULL, T_BOOL, T_NUM, T_STR, T_OBJ };

static void obj_release(Obj* o);
static inline void obj_retain(Obj* o);

struct Value {
    Tag tag;
    union {
        bool b;
        double n;
        Str* s;
        Obj* o;
    };
    Value() : tag(T_UNDEF), n(0) {}
    Value(double d) : tag(T_NUM), n(d) {}
    static Value boolean(bool v)
    {
        Value r;
        r.tag = T_BOOL;
        r.b = v;
        return r;
    }
    static Value null()
    {
        Value r;
        r.tag = T_NULL;
        return r;
    }
    static Value str(const U16& u)
    {
        Value r;
        r.tag = T_STR;
        r.s = new Str();
        r.s->s = u;
        r.s->rc = 1;
        return r;
    }
    static Value obj(Obj* o)
    {
        Value r;
        r.tag = T_OBJ;
        r.o = o;
        obj_retain(o);
        return r;
    }
    Value(const Value& v) : tag(v.tag), n(0)
    {
        copy_payload(v);
        retain();
    }
    Value(Value&& v) noexcept : tag(v.tag), n(0)
    {
        copy_payload(v);
        v.tag = T_UNDEF;
    }
    Value& operator=(const Value& v)
    {
        if (this != &v) {
            Value tmp(v);
            swap(tmp);
        }
        return *this;
    }
    Value& operator=(Value&& v) noexcept
    {
        if (this != &v) {
            release();
            tag = v.tag;
            copy_payload(v);
            v.tag = T_UNDEF;
        }
        return *this;
    }
    ~Value() { release(); }
    void swap(Value& o)
    {
        Tag t = tag;
        double pn = n;
        Str* ps = s;
        Obj* po = this->o;
        bool pb = b;
        tag = o.tag;
        copy_payload(o);
        o.tag = t;
        switch (t) {
            case T_BOOL: o.b = pb; break;
            case T_NUM: o.n = pn; break;
            case T_STR: o.s = ps; break;
            case T_OBJ: o.o = po; break;
            default: break;
        }
    }
    bool is_undef() const { return tag == T_UNDEF; }
    bool is_num() const { return tag == T_NUM; }
    bool is_obj() const { return tag == T_OBJ; }
    bool is_str() const { return tag == T_STR; }

private:
    void copy_payload(const Value& v)
    {
        switch (v.tag) {
            case T_BOOL: b = v.b; break;
            case T_NUM: n = v.n; break;
            case T_STR: s = v.s; break;
            case T_OBJ: o = v.o; break;
            default: break;
        }
    }
    void retain()
    {
        if (tag == T_STR)
            s->rc++;
        else if (tag == T_OBJ)
            obj_retain(o);
    }
    void release()
    {
        if (tag == T_STR) {
            if (--s->rc == 0) delete s;
        } else if (tag == T_OBJ) {
            obj_release(o);
        }
        tag = T_UNDEF;
    }
};

struct JsThrow {
    Value v;
};

// ---- atoms ------------------------------------------------------------------------------------------
struct AtomTable {
    std::unordered_map<std::string, Atom> map;
    std::vector<std::string> names;
    Atom get(const std::string& s)
    {
        auto it = map.find(s);
        if (it != map.end()) return it->second;
        Atom a = (Atom)names.size();
        names.push_back(s);
        map.emplace(s, a);
        return a;
    }
};
static AtomTable g_atoms;
static inline Atom A(const char* s) { return g_atoms.get(s); }

static std::string to_utf8(const U16& u)
{
    std::string r;
    for (size_t i = 0; i < u.size(); ++i) {
        uint32_t c = u[i];
        if (c >= 0xD800 && c < 0xDC00 && i + 1 < u.size() && u[i + 1] >= 0xDC00 && u[i + 1] < 0xE000) {
            c = 0x10000 + ((c - 0xD800) << 10) + (u[i + 1] - 0xDC00);
            ++i;
        }
        if (c < 0x80)
            r += (char)c;
        else if (c < 0x800) {
            r += (char)(0xC0 | (c >> 6));
            r += (char)(0x80 | (c & 63));
        } else if (c < 0x10000) {
            r += (char)(0xE0 | (c >> 12));
            r += (char)(0x80 | ((c >> 6) & 63));
            r += (char)(0x80 | (c & 63));
        } else {
            r += (char)(0xF0 | (c >> 18));
            r += (char)(0x80 | ((c >> 12) & 63));
            r += (char)(0x80 | ((c >> 6) & 63));
            r += (char)(0x80 | (c & 63));
        }
    }
    return r;
}
static U16 from_utf8(const std::string& s)
{
    U16 r;
    for (size_t i = 0; i < s.size();) {
        uint32_t c = (uint8_t)s[i];
        int extra = 0;
        if (c >= 0xF0) {
            c &= 7;
            extra = 3;
        } else if (c >= 0xE0) {
            c &= 15;
            extra = 2;
        } else if (c >= 0xC0) {
            c &= 31;
            extra = 1;
        }
        ++i;
        for (int k = 0; k < extra && i < s.size(); ++k, ++i) c = (c << 6) | ((uint8_t)s[i] & 63);
        if (c >= 0x10000) {
            c -= 0x10000;
            r += (char16_t)(0xD800 + (c >> 10));
            r += (char16_t)(0xDC00 + (c & 0x3FF));
        } else
            r += (char16_t)c;
    }
    return r;
}
static U16 ascii(const char* s)
{
    U16 r;
    while (*s) r += (char16_t)(uint8_t)*s++;
    return r;
}

// =====================================================================================================
// Objects
// =====================================================================================================
struct FuncNode;
struct Env;
struct Interp;

struct Prop {
    Value v;        // data property
    Obj* getter;    // accessor (not retained: class prototypes live for the whole run)
    Obj* setter;
    Prop() : getter(nullptr), setter(nullptr) {}
};

// open-addressing table Atom -> Prop, insertion order kept for Object.keys / assign
struct PropMap {
    std::vector<Atom> keys;     // insertion order (erased ones stay as ~0u)
    std::vector<Prop> vals;
    std::vector<int32_t> idx;   // hash slots -> index into keys / vals, -1 empty
    Prop* find(Atom a)
    {
        if (keys.size() <= 8) {
            for (size_t i = 0; i < keys.size(); ++i)
                if (keys[i] == a) return &vals[i];
            return nullptr;
        }
        if (idx.empty()) rehash();
        size_t mask = idx.size() - 1, h = (a * 2654435761u) & mask;
        while (idx[h] >= 0) {
            if (keys[idx[h]] == a) return &vals[idx[h]];
            h = (h + 1) & mask;
        }
        return nullptr;
    }
    Prop* insert(Atom a)
    {
        if (Prop* p = find(a)) return p;
        keys.push_back(a);
        vals.emplace_back();
        if (keys.size() > 8) {
            if (idx.empty() || keys.size() * 2 > idx.size())
                rehash();
            else {
                size_t mask = idx.size() - 1, h = (a * 2654435761u) & mask;
                while (idx[h] >= 0) h = (h + 1) & mask;
                idx[h] = (int32_t)keys.size() - 1;
            }
        }
        return &vals.back();
    }
    void erase(Atom a)
    {
        for (size_t i = 0; i < keys.size(); ++i)
            if (keys[i] == a) {
                keys.erase(keys.begin() + i);
                vals.erase(vals.begin() + i);
                idx.clear();
                return;
            }
    }
    void rehash()
    {
        size_t cap = 16;
        while (cap < keys.size() * 4) cap <<= 1;
        idx.assign(cap, -1);
        for (size_t i = 0; i < keys.size(); ++i) {
            size_t mask = cap - 1, h = (keys[i] * 2654435761u) & mask;
            while (idx[h] >= 0) h = (h + 1) & mask;
            idx[h] = (int32_t)i;
        }
    }
};

enum ObjKind : uint8_t { O_PLAIN, O_ARRAY, O_FUNC, O_TYPED, O_BUFFER, O_DATE };
enum ElemKind : uint8_t { E_U8, E_U16, E_U32, E_I8, E_I16, E_I32 };
static const int kElemSize[] = {1, 2, 4, 1, 2, 4};

typedef Value (*NativeFn)(Interp&, const Value& self, const Value* args, int argc);

struct Obj {
    uint32_t rc = 0;
    ObjKind kind;
    bool has_accessors = false;
    Obj* proto = nullptr;  // retained
    PropMap props;
    std::unordered_map<uint32_t, Value>* iprops = nullptr;  // integer keys of plain objects
    explicit Obj(ObjKind k) : kind(k) {}
    virtual ~Obj()
    {
        delete iprops;
        if (proto) obj_release(proto);
    }
    void set_proto(Obj* p)
    {
        if (p) obj_retain(p);
        if (proto) obj_release(proto);
        proto = p;
    }
};
static inline void obj_retain(Obj* o) { o->rc++; }
static void obj_release(Obj* o)
{
    if (--o->rc == 0) delete o;
}

struct ArrayObj : Obj {
    std::vector<Value> el;
    size_t head = 0;  // shift() moves the head instead of the elements
    ArrayObj() : Obj(O_ARRAY) {}
    size_t size() const { return el.size() - head; }
    Value& at(size_t i) { return el[head + i]; }
};

struct BufferObj : Obj {
    std::vector<uint8_t> data;
    BufferObj() : Obj(O_BUFFER) {}
};

struct TypedObj : Obj {
    BufferObj* buf = nullptr;  // retained
    size_t off = 0, len = 0;   // byte offset, element count
    ElemKind ek = E_U8;
    TypedObj() : Obj(O_TYPED) {}
    ~TypedObj() override
    {
        if (buf) obj_release(buf);
    }
    uint8_t* ptr() const { return buf->data.data() + off; }
};

struct DateObj : Obj {
    double ms = 0;
    DateObj() : Obj(O_DATE) {}
};

struct FuncObj : Obj {
    FuncNode* node = nullptr;
    Env* env = nullptr;  // retained
    NativeFn native = nullptr;
    int ctor_kind = 0;   // natives: which built-in constructor (typed array element kind etc.)
    bool is_class = false;
    std::string name;
    FuncObj() : Obj(O_FUNC) {}
    ~FuncObj() override;
};

struct Env {
    uint32_t rc = 0;
    Env* parent = nullptr;  // retained
    Value this_val;
    std::vector<Value> slots;
};
static inline void env_retain(Env* e)
{
    if (e) e->rc++;
}
static void env_release(Env* e)
{
    while (e && --e->rc == 0) {
        Env* p = e->parent;
        delete e;
        e = p;
    }
}
FuncObj::~FuncObj() { env_release(env); }

// =====================================================================================================
// Lexer
// =====================================================================================================
enum TokKind { TK_EOF, TK_NUM, TK_STR, TK_IDENT, TK_PUNCT, TK_KEYWORD };
struct Token {
    TokKind kind = TK_EOF;
    std::string text;  // identifier / punctuator / keyword
    U16 str;           // string literal value
    double num = 0;
    int line = 0;
    bool nl_before = false;  // a line terminator precedes the token
};

static const char* const kKeywords[] = {"var",   "let",    "const",  "function", "return", "if",     "else",       "for",
                                        "while", "do",     "break",  "continue", "new",    "this",   "typeof",     "void",
                                        "null",  "true",   "false",  "class",    "static", "throw",  "switch",     "case",
                                        "default", "instanceof", "in", "delete", "try",    "catch",  "finally",    "get",
                                        "set",   "of",     nullptr};
// get / set / static / of are contextual: the lexer reports them as identifiers
static bool is_keyword(const std::string& s)
{
    static const char* const hard[] = {"var",   "let",    "const", "function", "return", "if",    "else",    "for",
                                       "while", "do",     "break", "continue", "new",    "this",  "typeof",  "void",
                                       "null",  "true",   "false", "class",    "throw",  "switch", "case",   "default",
                                       "instanceof", "in", "delete", "try",    "catch",  "finally", nullptr};
    for (int i = 0; hard[i]; ++i)
        if (s == hard[i]) return true;
    return false;
}

struct Lexer {
    std::string src;
    size_t pos = 0;
    int line = 1;
    std::string file;
    [[noreturn]] void fail(const std::string& msg)
    {
        throw std::runtime_error(file + ":" + std::to_string(line) + ": " + msg);
    }
    Token next()
    {
        Token t;
        // whitespace and comments
        for (;;) {
            if (pos >= src.size()) break;
            char c = src[pos];
            if (c == '\n') {
                ++line;
                ++pos;
                t.nl_before = true;
            } else if (c == ' ' || c == '\t' || c == '\r') {
                ++pos;
            } else if (c == '/' && pos + 1 < src.size() && src[pos + 1] == '/') {
                while (pos < src.size() && src[pos] != '\n') ++pos;
            } else if (c == '/' && pos + 1 < src.size() && src[pos + 1] == '*') {
                pos += 2;
                while (pos + 1 < src.size() && !(src[pos] == '*' && src[pos + 1] == '/')) {
                    if (src[pos] == '\n') {
                        ++line;
                        t.nl_before = true;
                    }
                    ++pos;
                }
                pos += 2;
            } else
                break;
        }
        t.line = line;
        if (pos >= src.size()) {
            t.kind = TK_EOF;
            return t;
        }
        char c = src[pos];
        if (isdigit((unsigned char)c) || (c == '.' && pos + 1 < src.size() && isdigit((unsigned char)src[pos + 1]))) {
            t.kind = TK_NUM;
            if (c == '0' && pos + 1 < src.size() && (src[pos + 1] == 'x' || src[pos + 1] == 'X')) {
                pos += 2;
                double v = 0;
                while (pos < src.size() && isxdigit((unsigned char)src[pos])) {
                    char d = src[pos++];
                    v = v * 16 + (isdigit((unsigned char)d) ? d - '0' : (tolower(d) - 'a' + 10));
                }
                t.num = v;
            } else if (c == '0' && pos + 1 < src.size() && (src[pos + 1] == 'b' || src[pos + 1] == 'B')) {
                pos += 2;
                double v = 0;
                while (pos < src.size() && (src[pos] == '0' || src[pos] == '1')) v = v * 2 + (src[pos++] - '0');
                t.num = v;
            } else {
                size_t st = pos;
                while (pos < src.size() && (isdigit((unsigned char)src[pos]) || src[pos] == '.')) ++pos;
                if (pos < src.size() && (src[pos] == 'e' || src[pos] == 'E')) {
                    ++pos;
                    if (pos < src.size() && (src[pos] == '+' || src[pos] == '-')) ++pos;
                    while (pos < src.size() && isdigit((unsigned char)src[pos])) ++pos;
                }
                t.num = strtod(src.substr(st, pos - st).c_str(), nullptr);
            }
            return t;
        }
        if (isalpha((unsigned char)c) || c == '_' || c == '$') {
            size_t st = pos;
            while (pos < src.size() && (isalnum((unsigned char)src[pos]) || src[pos] == '_' || src[pos] == '$')) ++pos;
            t.text = src.substr(st, pos - st);
            t.kind = is_keyword(t.text) ? TK_KEYWORD : TK_IDENT;
            return t;
        }
        if (c == '"' || c == '\'') {
            char q = c;
            ++pos;
            std::string raw;
            while (pos < src.size() && src[pos] != q) {
                char d = src[pos++];
                if (d == '\\') {
                    char e = src[pos++];
                    switch (e) {
                        case 'n': raw += '\n'; break;
                        case 't': raw += '\t'; break;
                        case 'r': raw += '\r'; break;
                        case 'b': raw += '\b'; break;
                        case 'f': raw += '\f'; break;
                        case 'v': raw += '\v'; break;
                        case '0': raw += '\0'; break;
                        case 'x': {
                            int v = (int)strtol(src.substr(pos, 2).c_str(), nullptr, 16);
                            pos += 2;
                            raw += to_utf8(U16(1, (char16_t)v));
                            break;
                        }
                        case 'u': {
                            int v = (int)strtol(src.substr(pos, 4).c_str(), nullptr, 16);
                            pos += 4;
                            raw += to_utf8(U16(1, (char16_t)v));
                            break;
                        }
                        case '\n': ++line; break;
                        default: raw += e;
                    }
                } else
                    raw += d;
            }
            if (pos >= src.size()) fail("unterminated string");
            ++pos;
            t.kind = TK_STR;
            t.str = from_utf8(raw);
            return t;
        }
        static const char* const puncts[] = {">>>=", "...", "===", "!==", "**=", "<<=", ">>=", ">>>", "&&=", "||=", "??=",
                                             "=>",   "==",  "!=",  "<=",  ">=",  "&&",  "||",  "??",  "++",  "--",  "+=",
                                             "-=",   "*=",  "/=",  "%=",  "&=",  "|=",  "^=",  "<<",  ">>",  "**",  "?.",
                                             nullptr};
        for (int i = 0; puncts[i]; ++i) {
            size_t L = strlen(puncts[i]);
            if (src.compare(pos, L, puncts[i]) == 0) {
                t.kind = TK_PUNCT;
                t.text = puncts[i];
                pos += L;
                return t;
            }
        }
        t.kind = TK_PUNCT;
        t.text = std::string(1, c);
        ++pos;
        return t;
    }
};

// =====================================================================================================
// AST
// =====================================================================================================
enum NodeKind : uint8_t {
    // expressions
    N_NUM, N_STR, N_IDENT, N_THIS, N_NULL, N_UNDEF, N_TRUE, N_FALSE, N_ARRAY, N_OBJECT, N_FUNC, N_CLASS,
    N_MEMBER, N_INDEX, N_CALL, N_NEW, N_UNARY, N_UPDATE, N_BINARY, N_LOGICAL, N_ASSIGN, N_COND, N_SEQ,
    // statements
    N_VAR, N_EXPR, N_BLOCK, N_IF, N_FOR, N_FORIN, N_WHILE, N_DOWHILE, N_RETURN, N_BREAK, N_CONTINUE, N_THROW,
    N_SWITCH, N_TRY, N_EMPTY, N_FUNCDECL
};

enum Op : uint8_t {
    OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_MOD, OP_POW, OP_SHL, OP_SHR, OP_USHR, OP_AND, OP_OR, OP_XOR,
    OP_LT, OP_GT, OP_LE, OP_GE, OP_EQ, OP_NE, OP_SEQ, OP_SNE, OP_INSTANCEOF, OP_IN,
    OP_LAND, OP_LOR, OP_NULLISH,
    OP_NOT, OP_BITNOT, OP_NEG, OP_PLUS, OP_TYPEOF, OP_VOID, OP_DELETE,
    OP_NONE
};

struct Node;
struct PropDef {
    Atom key = 0;
    bool is_index = false;  // numeric key
    uint32_t index = 0;
    Node* computed = nullptr;
    Node* value = nullptr;  // expression or function
    int accessor = 0;       // 1 getter, 2 setter
    bool is_static = false;
};
struct CaseDef {
    Node* test = nullptr;  // nullptr = default
    std::vector<Node*> body;
};

struct Node {
    NodeKind kind;
    Op op = OP_NONE;
    int line = 0;
    double num = 0;
    Value strv;                 // N_STR
    Atom atom = 0;              // identifier / member name
    int hops = -1, slot = -1;   // resolved variable (-1: global)
    bool prefix = false;        // N_UPDATE
    bool for_of = false;        // N_FORIN
    Node *a = nullptr, *b = nullptr, *c = nullptr, *d = nullptr;
    std::vector<Node*> list;    // args, elements, statements, declarators (pairs ident/init)
    std::vector<PropDef> props; // object literal / class body
    std::vector<CaseDef> cases;
    FuncNode* fn = nullptr;
    explicit Node(NodeKind k) : kind(k) {}
};

struct FuncNode {
    std::string name;
    std::vector<Atom> locals;       // params first, then every var / function / class name of the body
    std::vector<Node*> param_defaults;  // per param, nullptr when none
    int n_params = 0;
    bool is_arrow = false;
    bool expr_body = false;
    Node* body = nullptr;           // N_BLOCK, or an expression for arrows with expr_body
    FuncNode* parent = nullptr;
    std::vector<Node*> idents;      // identifier nodes to resolve
    std::vector<Node*> func_decls;  // hoisted function declarations (N_FUNCDECL)
    Atom self_name = 0;             // named function expression
    bool has_self = false;
    int slot_of(Atom a) const
    {
        for (size_t i = 0; i < locals.size(); ++i)
            if (locals[i] == a) return (int)i;
        return -1;
    }
    int declare(Atom a)
    {
        int s = slot_of(a);
        if (s >= 0) return s;
        locals.push_back(a);
        return (int)locals.size() - 1;
    }
};

// =====================================================================================================
// Parser
// =====================================================================================================
struct Parser {
    Lexer lx;
    Token tok, peeked;
    bool has_peek = false;
    FuncNode* cur = nullptr;

    void advance()
    {
        if (has_peek) {
            tok = peeked;
            has_peek = false;
        } else
            tok = lx.next();
    }
    const Token& peek()
    {
        if (!has_peek) {
            peeked = lx.next();
            has_peek = true;
        }
        return peeked;
    }
    [[noreturn]] void fail(const std::string& m)
    {
        throw std::runtime_error(lx.file + ":" + std::to_string(tok.line) + ": " + m + " (at '" + tok.text + "')");
    }
    bool is_p(const char* p) const { return tok.kind == TK_PUNCT && tok.text == p; }
    bool is_k(const char* k) const { return tok.kind == TK_KEYWORD && tok.text == k; }
    bool is_id(const char* k) const { return tok.kind == TK_IDENT && tok.text == k; }
    void expect_p(const char* p)
    {
        if (!is_p(p)) fail(std::string("expected '") + p + "'");
        advance();
    }
    bool accept_p(const char* p)
    {
        if (is_p(p)) {
            advance();
            return true;
        }
        return false;
    }
    void semicolon()
    {
        if (accept_p(";")) return;
        if (is_p("}") || tok.kind == TK_EOF || tok.nl_before) return;  // automatic semicolon insertion
        fail("expected ';'");
    }
    Node* mk(NodeKind k)
    {
        Node* n = new Node(k);
        n->line = tok.line;
        return n;
    }
    std::string ident_name()
    {
        // identifiers and keywords are both fine as property names
        if (tok.kind != TK_IDENT && tok.kind != TK_KEYWORD) fail("expected a name");
        std::string s = tok.text;
        advance();
        return s;
    }

    FuncNode* parse_program()
    {
        FuncNode* f = new FuncNode();
        f->name = "<program>";
        cur = f;
        advance();
        Node* blk = mk(N_BLOCK);
        while (tok.kind != TK_EOF) blk->list.push_back(statement());
        f->body = blk;
        return f;
    }

    // ---- functions
    // parses "(params) { body }" (after an optional name) into a FuncNode
    FuncNode* function_rest(const std::string& name, bool arrow_params_done = false)
    {
        (void)arrow_params_done;
        FuncNode* f = new FuncNode();
        f->name = name;
        f->parent = cur;
        cur = f;
        expect_p("(");
        params(f);
        expect_p(")");
        f->body = block();
        cur = f->parent;
        return f;
    }
    void params(FuncNode* f)
    {
        while (!is_p(")")) {
            if (tok.kind != TK_IDENT) fail("expected a parameter name");
            f->locals.push_back(g_atoms.get(tok.text));
            advance();
            Node* def = nullptr;
            if (accept_p("=")) def = assignment();
            f->param_defaults.push_back(def);
            if (!accept_p(",")) break;
        }
        f->n_params = (int)f->locals.size();
    }

    Node* block()
    {
        Node* b = mk(N_BLOCK);
        expect_p("{");
        while (!is_p("}")) {
            if (tok.kind == TK_EOF) fail("unterminated block");
            b->list.push_back(statement());
        }
        advance();
        return b;
    }

    // ---- statements
    Node* statement()
    {
        if (tok.kind == TK_PUNCT) {
            if (is_p("{")) return block();
            if (is_p(";")) {
                advance();
                return mk(N_EMPTY);
            }
        }
        if (tok.kind == TK_KEYWORD) {
            const std::string& k = tok.text;
            if (k == "var" || k == "let" || k == "const") {
                Node* n = var_decl();
                semicolon();
                return n;
            }
            if (k == "function") {
                Node* n = mk(N_FUNCDECL);
                advance();
                std::string name = ident_name();
                n->atom = g_atoms.get(name);
                cur->declare(n->atom);
                Node* id = mk(N_IDENT);
                id->atom = n->atom;
                cur->idents.push_back(id);
                n->a = id;
                n->fn = function_rest(name);
                cur->func_decls.push_back(n);
                return n;
            }
            if (k == "class") {
                // class declaration = var Name = class Name {...}
                Node* cls = class_expr();
                Node* n = mk(N_VAR);
                Node* id = mk(N_IDENT);
                id->atom = cls->atom;
                cur->declare(id->atom);
                cur->idents.push_back(id);
                n->list.push_back(id);
                n->list.push_back(cls);
                return n;
            }
            if (k == "return") {
                Node* n = mk(N_RETURN);
                advance();
                if (!is_p(";") && !is_p("}") && !tok.nl_before && tok.kind != TK_EOF) n->a = expression();
                semicolon();
                return n;
            }
            if (k == "if") {
                Node* n = mk(N_IF);
                advance();
                expect_p("(");
                n->a = expression();
                expect_p(")");
                n->b = statement();
                if (is_k("else")) {
                    advance();
                    n->c = statement();
                }
                return n;
            }
            if (k == "for") return for_stmt();
            if (k == "while") {
                Node* n = mk(N_WHILE);
                advance();
                expect_p("(");
                n->a = expression();
                expect_p(")");
                n->b = statement();
                return n;
            }
            if (k == "do") {
                Node* n = mk(N_DOWHILE);
                advance();
                n->b = statement();
                if (!is_k("while")) fail("expected 'while'");
                advance();
                expect_p("(");
                n->a = expression();
                expect_p(")");
                accept_p(";");
                return n;
            }
            if (k == "break" || k == "continue") {
                Node* n = mk(k == "break" ? N_BREAK : N_CONTINUE);
                advance();
                semicolon();
                return n;
            }
            if (k == "throw") {
                Node* n = mk(N_THROW);
                advance();
                n->a = expression();
                semicolon();
                return n;
            }
            if (k == "switch") {
                Node* n = mk(N_SWITCH);
                advance();
                expect_p("(");
                n->a = expression();
                expect_p(")");
                expect_p("{");
                while (!is_p("}")) {
                    CaseDef cd;
                    if (is_k("case")) {
                        advance();
                        cd.test = expression();
                    } else if (is_k("default")) {
                        advance();
                    } else
                        fail("expected 'case' or 'default'");
                    expect_p(":");
                    while (!is_p("}") && !is_k("case") && !is_k("default")) cd.body.push_back(statement());
                    n->cases.push_back(cd);
                }
                advance();
                return n;
            }
            if (k == "try") {
                Node* n = mk(N_TRY);
                advance();
                n->a = block();
                if (is_k("catch")) {
                    advance();
                    if (accept_p("(")) {
                        Node* id = mk(N_IDENT);
                        id->atom = g_atoms.get(ident_name());
                        cur->declare(id->atom);
                        cur->idents.push_back(id);
                        n->d = id;
                        expect_p(")");
                    }
                    n->b = block();
                }
                if (is_k("finally")) {
                    advance();
                    n->c = block();
                }
                return n;
            }
        }
        Node* n = mk(N_EXPR);
        n->a = expression();
        semicolon();
        return n;
    }

    Node* var_decl()
    {
        Node* n = mk(N_VAR);
        advance();  // var / let / const: all function scoped here (the bundle has no shadowing block bindings)
        for (;;) {
            if (tok.kind != TK_IDENT) fail("expected a variable name");
            Node* id = mk(N_IDENT);
            id->atom = g_atoms.get(tok.text);
            cur->declare(id->atom);
            cur->idents.push_back(id);
            advance();
            Node* init = nullptr;
            if (accept_p("=")) init = assignment();
            n->list.push_back(id);
            n->list.push_back(init);
            if (!accept_p(",")) break;
        }
        return n;
    }

    Node* for_stmt()
    {
        advance();
        expect_p("(");
        Node* init = nullptr;
        if (is_p(";")) {
        } else if (is_k("var") || is_k("let") || is_k("const")) {
            init = var_decl();
        } else {
            init = mk(N_EXPR);
            init->a = expression_no_in();
        }
        if (is_k("in") || is_id("of")) {
            Node* n = mk(N_FORIN);
            n->for_of = is_id("of");
            advance();
            // target: the declared variable or an assignable expression
            n->a = init->kind == N_VAR ? init->list[0] : init->a;
            n->b = expression();
            expect_p(")");
            n->c = statement();
            return n;
        }
        Node* n = mk(N_FOR);
        n->a = init;
        expect_p(";");
        if (!is_p(";")) n->b = expression();
        expect_p(";");
        if (!is_p(")")) n->c = expression();
        expect_p(")");
        n->d = statement();
        return n;
    }

    // ---- expressions
    bool no_in = false;
    Node* expression_no_in()
    {
        bool old = no_in;
        no_in = true;
        Node* n = expression();
        no_in = old;
        return n;
    }
    Node* expression()
    {
        Node* n = assignment();
        if (is_p(",")) {
            Node* s = mk(N_SEQ);
            s->list.push_back(n);
            while (accept_p(",")) s->list.push_back(assignment());
            return s;
        }
        return n;
    }

    // is the parenthesis at the current token the start of arrow parameters?
    bool arrow_ahead()
    {
        // scan forward on a copy of the lexer: balanced parentheses followed by =>
        Lexer save = lx;
        Token st = tok, sp = peeked;
        bool hp = has_peek;
        int depth = 0;
        bool res = false;
        for (;;) {
            if (tok.kind == TK_EOF) break;
            if (is_p("(")) ++depth;
            if (is_p(")")) {
                --depth;
                if (depth == 0) {
                    advance();
                    res = is_p("=>");
                    break;
                }
            }
            advance();
        }
        lx = save;
        tok = st;
        peeked = sp;
        has_peek = hp;
        return res;
    }

    Node* arrow_function(bool parens)
    {
        Node* n = mk(N_FUNC);
        FuncNode* f = new FuncNode();
        f->is_arrow = true;
        f->name = "<arrow>";
        f->parent = cur;
        cur = f;
        if (parens) {
            expect_p("(");
            params(f);
            expect_p(")");
        } else {
            f->locals.push_back(g_atoms.get(tok.text));
            f->param_defaults.push_back(nullptr);
            f->n_params = 1;
            advance();
        }
        expect_p("=>");
        if (is_p("{")) {
            f->body = block();
        } else {
            f->expr_body = true;
            f->body = assignment();
        }
        cur = f->parent;
        n->fn = f;
        return n;
    }

    Node* assignment()
    {
        if (tok.kind == TK_IDENT && peek().kind == TK_PUNCT && peek().text == "=>") return arrow_function(false);
        if (is_p("(") && arrow_ahead()) return arrow_function(true);
        Node* lhs = conditional();
        if (tok.kind == TK_PUNCT) {
            static const struct {
                const char* p;
                Op op;
            } ops[] = {{"=", OP_NONE},   {"+=", OP_ADD},  {"-=", OP_SUB},   {"*=", OP_MUL},   {"/=", OP_DIV},
                       {"%=", OP_MOD},   {"<<=", OP_SHL}, {">>=", OP_SHR},  {">>>=", OP_USHR}, {"&=", OP_AND},
                       {"|=", OP_OR},    {"^=", OP_XOR},  {"**=", OP_POW},  {"&&=", OP_LAND},  {"||=", OP_LOR},
                       {"??=", OP_NULLISH}};
            for (auto& o : ops)
                if (tok.text == o.p) {
                    if (lhs->kind != N_IDENT && lhs->kind != N_MEMBER && lhs->kind != N_INDEX) fail("invalid assignment target");
                    Node* n = mk(N_ASSIGN);
                    n->op = o.op;
                    advance();
                    n->a = lhs;
                    n->b = assignment();
                    return n;
                }
        }
        return lhs;
    }

    Node* conditional()
    {
        Node* c = binary(0);
        if (is_p("?")) {
            Node* n = mk(N_COND);
            advance();
            n->a = c;
            bool old = no_in;
            no_in = false;
            n->b = assignment();
            no_in = old;
            expect_p(":");
            n->c = assignment();
            return n;
        }
        return c;
    }

    int binary_prec(Op& op, bool& logical)
    {
        logical = false;
        if (tok.kind == TK_KEYWORD) {
            if (tok.text == "instanceof") {
                op = OP_INSTANCEOF;
                return 9;
            }
            if (tok.text == "in" && !no_in) {
                op = OP_IN;
                return 9;
            }
            return -1;
        }
        if (tok.kind != TK_PUNCT) return -1;
        const std::string& t = tok.text;
        static const struct {
            const char* p;
            Op op;
            int prec;
            bool logical;
        } tab[] = {{"??", OP_NULLISH, 1, true}, {"||", OP_LOR, 2, true},  {"&&", OP_LAND, 3, true}, {"|", OP_OR, 4, false},
                   {"^", OP_XOR, 5, false},     {"&", OP_AND, 6, false},  {"==", OP_EQ, 7, false},  {"!=", OP_NE, 7, false},
                   {"===", OP_SEQ, 7, false},   {"!==", OP_SNE, 7, false}, {"<", OP_LT, 9, false},  {">", OP_GT, 9, false},
                   {"<=", OP_LE, 9, false},     {">=", OP_GE, 9, false},  {"<<", OP_SHL, 10, false}, {">>", OP_SHR, 10, false},
                   {">>>", OP_USHR, 10, false}, {"+", OP_ADD, 11, false}, {"-", OP_SUB, 11, false}, {"*", OP_MUL, 12, false},
                   {"/", OP_DIV, 12, false},    {"%", OP_MOD, 12, false}, {"**", OP_POW, 13, false}};
        for (auto& e : tab)
            if (t == e.p) {
                op = e.op;
                logical = e.logical;
                return e.prec;
            }
        return -1;
    }

    Node* binary(int min_prec)
    {
        Node* lhs = unary();
        for (;;) {
            Op op;
            bool logical;
            int prec = binary_prec(op, logical);
            if (prec < 0 || prec < min_prec) return lhs;
            Node* n = mk(logical ? N_LOGICAL : N_BINARY);
            n->op = op;
            advance();
            n->a = lhs;
            n->b = binary(op == OP_POW ? prec : prec + 1);  // ** is right associative
            lhs = n;
        }
    }

    Node* unary()
    {
        if (tok.kind == TK_PUNCT) {
            Op op = OP_NONE;
            if (tok.text == "!") op = OP_NOT;
            else if (tok.text == "~") op = OP_BITNOT;
            else if (tok.text == "-") op = OP_NEG;
            else if (tok.text == "+") op = OP_PLUS;
            if (op != OP_NONE) {
                Node* n = mk(N_UNARY);
                n->op = op;
                advance();
                n->a = unary();
                return n;
            }
            if (tok.text == "++" || tok.text == "--") {
                Node* n = mk(N_UPDATE);
                n->op = tok.text == "++" ? OP_ADD : OP_SUB;
                n->prefix = true;
                advance();
                n->a = unary();
                return n;
            }
        }
        if (tok.kind == TK_KEYWORD) {
            Op op = OP_NONE;
            if (tok.text == "typeof") op = OP_TYPEOF;
            else if (tok.text == "void") op = OP_VOID;
            else if (tok.text == "delete") op = OP_DELETE;
            if (op != OP_NONE) {
                Node* n = mk(N_UNARY);
                n->op = op;
                advance();
                n->a = unary();
                return n;
            }
        }
        Node* e = postfix();
        return e;
    }

    Node* postfix()
    {
        Node* e = call_member(primary());
        if (tok.kind == TK_PUNCT && (tok.text == "++" || tok.text == "--") && !tok.nl_before) {
            Node* n = mk(N_UPDATE);
            n->op = tok.text == "++" ? OP_ADD : OP_SUB;
            n->prefix = false;
            advance();
            n->a = e;
            return n;
        }
        return e;
    }

    void arguments(Node* n)
    {
        expect_p("(");
        bool old = no_in;
        no_in = false;
        while (!is_p(")")) {
            n->list.push_back(assignment());
            if (!accept_p(",")) break;
        }
        no_in = old;
        expect_p(")");
    }

    Node* call_member(Node* e)
    {
        for (;;) {
            if (is_p(".")) {
                advance();
                Node* n = mk(N_MEMBER);
                n->a = e;
                n->atom = g_atoms.get(ident_name());
                e = n;
            } else if (is_p("[")) {
                advance();
                Node* n = mk(N_INDEX);
                n->a = e;
                bool old = no_in;
                no_in = false;
                n->b = expression();
                no_in = old;
                expect_p("]");
                e = n;
            } else if (is_p("(")) {
                Node* n = mk(N_CALL);
                n->a = e;
                arguments(n);
                e = n;
            } else
                return e;
        }
    }

    Node* new_expr()
    {
        Node* n = mk(N_NEW);
        advance();  // new
        Node* callee;
        if (is_k("new"))
            callee = new_expr();
        else
            callee = primary();
        // member accesses bind tighter than the call of `new`
        for (;;) {
            if (is_p(".")) {
                advance();
                Node* m = mk(N_MEMBER);
                m->a = callee;
                m->atom = g_atoms.get(ident_name());
                callee = m;
            } else if (is_p("[")) {
                advance();
                Node* m = mk(N_INDEX);
                m->a = callee;
                m->b = expression();
                expect_p("]");
                callee = m;
            } else
                break;
        }
        n->a = callee;
        if (is_p("(")) arguments(n);
        return n;
    }

    Node* class_expr()
    {
        Node* n = mk(N_CLASS);
        advance();  // class
        std::string name;
        if (tok.kind == TK_IDENT) {
            name = tok.text;
            n->atom = g_atoms.get(name);
            // the inner name is visible in the body: a function-scoped binding of the enclosing function
            cur->declare(n->atom);
            Node* id = mk(N_IDENT);
            id->atom = n->atom;
            cur->idents.push_back(id);
            n->a = id;
            advance();
        }
        expect_p("{");
        while (!is_p("}")) {
            if (accept_p(";")) continue;
            PropDef pd;
            if (is_id("static") && !(peek().kind == TK_PUNCT && (peek().text == "(" || peek().text == "="))) {
                pd.is_static = true;
                advance();
            }
            if ((is_id("get") || is_id("set")) && !(peek().kind == TK_PUNCT && (peek().text == "(" || peek().text == "="))) {
                pd.accessor = is_id("get") ? 1 : 2;
                advance();
            }
            std::string key;
            if (tok.kind == TK_STR) {
                key = to_utf8(tok.str);
                advance();
            } else
                key = ident_name();
            pd.key = g_atoms.get(key);
            if (is_p("(")) {
                Node* fn = mk(N_FUNC);
                fn->fn = function_rest(key);
                pd.value = fn;
            } else {
                // field: name = expr;  (evaluated per instance for instance fields, once for static ones)
                if (accept_p("=")) {
                    // field initialisers run as little arrow functions so that `this` is the instance
                    Node* fnn = mk(N_FUNC);
                    FuncNode* f = new FuncNode();
                    f->is_arrow = false;
                    f->name = "<field " + key + ">";
                    f->parent = cur;
                    cur = f;
                    f->expr_body = true;
                    f->body = assignment();
                    cur = f->parent;
                    fnn->fn = f;
                    pd.value = fnn;
                }
                pd.accessor = pd.accessor ? pd.accessor : 3;  // 3 = field
                semicolon();
            }
            n->props.push_back(pd);
        }
        advance();
        (void)name;
        return n;
    }

    Node* primary()
    {
        switch (tok.kind) {
            case TK_NUM: {
                Node* n = mk(N_NUM);
                n->num = tok.num;
                advance();
                return n;
            }
            case TK_STR: {
                Node* n = mk(N_STR);
                n->strv = Value::str(tok.str);
                advance();
                return n;
            }
            case TK_IDENT: {
                Node* n = mk(N_IDENT);
                n->atom = g_atoms.get(tok.text);
                if (tok.text == "undefined") n->kind = N_UNDEF;
                else cur->idents.push_back(n);
                advance();
                return n;
            }
            case TK_KEYWORD: {
                const std::string& k = tok.text;
                if (k == "this") {
                    advance();
                    return mk(N_THIS);
                }
                if (k == "null") {
                    advance();
                    return mk(N_NULL);
                }
                if (k == "true") {
                    advance();
                    return mk(N_TRUE);
                }
                if (k == "false") {
                    advance();
                    return mk(N_FALSE);
                }
                if (k == "new") return new_expr();
                if (k == "class") return class_expr();
                if (k == "function") {
                    Node* n = mk(N_FUNC);
                    advance();
                    std::string name;
                    if (tok.kind == TK_IDENT) {
                        name = tok.text;
                        advance();
                    }
                    n->fn = function_rest(name);
                    if (!name.empty()) {
                        // named function expression: the name is a local of the function itself
                        n->fn->self_name = g_atoms.get(name);
                        n->fn->has_self = true;
                        n->fn->declare(n->fn->self_name);
                    }
                    return n;
                }
                fail("unexpected keyword");
            }
            case TK_PUNCT: {
                if (is_p("(")) {
                    advance();
                    bool old = no_in;
                    no_in = false;
                    Node* e = expression();
                    no_in = old;
                    expect_p(")");
                    return e;
                }
                if (is_p("[")) {
                    Node* n = mk(N_ARRAY);
                    advance();
                    while (!is_p("]")) {
                        if (is_p(",")) {
                            advance();
                            n->list.push_back(mk(N_UNDEF));
                            continue;
                        }
                        n->list.push_back(assignment());
                        if (!accept_p(",")) break;
                    }
                    expect_p("]");
                    return n;
                }
                if (is_p("{")) return object_literal();
                fail("unexpected token");
            }
            default: fail("unexpected end of input");
        }
    }

    Node* object_literal()
    {
        Node* n = mk(N_OBJECT);
        advance();
        while (!is_p("}")) {
            PropDef pd;
            if ((is_id("get") || is_id("set")) &&
                !(peek().kind == TK_PUNCT && (peek().text == "(" || peek().text == ":" || peek().text == "," || peek().text == "}"))) {
                pd.accessor = is_id("get") ? 1 : 2;
                advance();
            }
            bool was_ident = tok.kind == TK_IDENT;
            std::string key;
            if (tok.kind == TK_NUM) {
                double v = tok.num;
                if (v >= 0 && v < 4294967295.0 && v == floor(v)) {
                    pd.is_index = true;
                    pd.index = (uint32_t)v;
                } else {
                    char buf[64];
                    snprintf(buf, sizeof buf, "%.17g", v);
                    key = buf;
                }
                advance();
            } else if (tok.kind == TK_STR) {
                key = to_utf8(tok.str);
                advance();
            } else if (is_p("[")) {
                advance();
                pd.computed = assignment();
                expect_p("]");
            } else
                key = ident_name();
            if (!pd.is_index && !pd.computed) pd.key = g_atoms.get(key);
            if (is_p("(")) {
                Node* fn = mk(N_FUNC);
                fn->fn = function_rest(key);
                pd.value = fn;
            } else if (accept_p(":")) {
                pd.value = assignment();
            } else {
                // shorthand { name }
                if (!was_ident) fail("expected ':'");
                Node* id = mk(N_IDENT);
                id->atom = pd.key;
                cur->idents.push_back(id);
                pd.value = id;
            }
            n->props.push_back(pd);
            if (!accept_p(",")) break;
        }
        expect_p("}");
        return n;
    }
};

// resolve identifiers to (hops, slot); everything else is a global
static void resolve(FuncNode* f)
{
    for (Node* id : f->idents) {
        int hops = 0;
        for (FuncNode* s = f; s; s = s->parent, ++hops) {
            int sl = s->slot_of(id->atom);
            if (sl >= 0) {
                id->hops = hops;
                id->slot = sl;
                break;
            }
        }
    }
}
static void resolve_tree(Node* n);
static void resolve_func(FuncNode* f)
{
    resolve(f);
    resolve_tree(f->body);
    for (Node* d : f->param_defaults)
        if (d) resolve_tree(d);
}
static void resolve_tree(Node* n)
{
    if (!n) return;
    if (n->fn) resolve_func(n->fn);
    resolve_tree(n->a);
    resolve_tree(n->b);
    resolve_tree(n->c);
    resolve_tree(n->d);
    for (Node* c : n->list) resolve_tree(c);
    for (auto& p : n->props) {
        resolve_tree(p.value);
        resolve_tree(p.computed);
    }
    for (auto& c : n->cases) {
        resolve_tree(c.test);
        for (Node* s : c.body) resolve_tree(s);
    }
}

// =====================================================================================================
// Interpreter
// =====================================================================================================
enum Completion { C_NORMAL, C_BREAK, C_CONTINUE, C_RETURN };

struct Interp {
    Obj* global = nullptr;
    Obj *object_proto = nullptr, *function_proto = nullptr, *array_proto = nullptr, *error_proto = nullptr,
        *date_proto = nullptr, *buffer_proto = nullptr, *string_proto = nullptr, *number_proto = nullptr;
    Obj* typed_proto[6] = {nullptr};
    FuncObj* typed_ctor[6] = {nullptr};
    FuncObj* error_ctor = nullptr;
    std::vector<std::string> args;
    Value ret;  // value of the return statement in flight
    int depth = 0;

    Atom a_length, a_prototype, a_constructor, a_message, a_name, a_buffer, a_byteOffset, a_byteLength,
        a_BYTES_PER_ELEMENT, a_stack;

    // ---- helpers
    [[noreturn]] void throw_error(const std::string& msg, const char* kind = "Error");
    Value make_error(const std::string& msg, const char* kind);
    FuncObj* make_native(const char* name, NativeFn fn, int ctor_kind = 0);
    void def_native(Obj* target, const char* name, NativeFn fn, int ctor_kind = 0);
    void setup();

    // conversions
    double to_number(const Value& v);
    int32_t to_int32(const Value& v) { return to_int32(to_number(v)); }
    static int32_t to_int32(double d)
    {
        if (d >= -2147483648.0 && d <= 2147483647.0) return (int32_t)d;  // fast path (truncation)
        if (std::isnan(d) || std::isinf(d)) return 0;
        double t = std::trunc(d);
        double m = std::fmod(t, 4294967296.0);
        if (m < 0) m += 4294967296.0;
        return (int32_t)(uint32_t)m;
    }
    static uint32_t to_uint32(double d) { return (uint32_t)to_int32(d); }
    bool to_bool(const Value& v);
    U16 to_string(const Value& v);
    static U16 num_to_string(double d, int radix = 10);
    Value to_primitive(const Value& v);
    U16 type_of(const Value& v);

    // property access
    Value get_prop(const Value& obj, Atom a);
    Value get_prop_obj(Obj* o, Atom a, const Value& self);
    void set_prop(const Value& obj, Atom a, const Value& v);
    Value get_index(const Value& obj, const Value& key);
    void set_index(const Value& obj, const Value& key, const Value& v);
    bool key_to_index(const Value& key, uint32_t& idx);
    Atom key_to_atom(const Value& key);
    bool has_property(Obj* o, const Value& key);

    // typed arrays
    static Value typed_get(TypedObj* t, size_t i);
    void typed_set(TypedObj* t, size_t i, const Value& v);
    TypedObj* new_typed(ElemKind ek, size_t len);
    TypedObj* new_typed_on(ElemKind ek, BufferObj* b, size_t off, size_t len);
    ArrayObj* new_array();
    Obj* new_object();

    // evaluation
    Value eval(Node* n, Env* env);
    Completion exec(Node* n, Env* env);
    Completion exec_list(const std::vector<Node*>& l, Env* env);
    Value call(const Value& fn, const Value& self, const Value* args, int argc);
    Value call_function(FuncObj* f, const Value& self, const Value* args, int argc);
    Value construct(const Value& fn, const Value* args, int argc);
    Value make_closure(FuncNode* fn, Env* env);
    Value eval_class(Node* n, Env* env);
    Value binary_op(Op op, const Value& a, const Value& b);
    bool strict_equals(const Value& a, const Value& b);
    bool loose_equals(const Value& a, const Value& b);
    bool instance_of(const Value& v, const Value& ctor);
    Value& var_ref(Node* id, Env* env, bool& is_global);
    void assign_to(Node* target, const Value& v, Env* env);
    Env* this_env(Env* env);
};

static std::string narrow(const U16& u) { return to_utf8(u); }

Value Interp::make_error(const std::string& msg, const char* kind)
{
    Obj* e = new Obj(O_PLAIN);
    e->set_proto(error_proto);
    Value v = Value::obj(e);
    e->props.insert(a_message)->v = Value::str(from_utf8(msg));
    e->props.insert(a_name)->v = Value::str(ascii(kind));
    return v;
}
void Interp::throw_error(const std::string& msg, const char* kind) { throw JsThrow{make_error(msg, kind)}; }

double Interp::to_number(const Value& v)
{
    switch (v.tag) {
        case T_NUM: return v.n;
        case T_UNDEF: return NAN;
        case T_NULL: return 0;
        case T_BOOL: return v.b ? 1 : 0;
        case T_STR: {
            std::string s = narrow(v.s->s);
            size_t a = 0, b = s.size();
            while (a < b && isspace((unsigned char)s[a])) ++a;
            while (b > a && isspace((unsigned char)s[b - 1])) --b;
            if (a == b) return 0;
            s = s.substr(a, b - a);
            if (s.size() > 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) {
                char* end;
                double r = (double)strtoull(s.c_str() + 2, &end, 16);
                return *end ? NAN : r;
            }
            if (s == "Infinity" || s == "+Infinity") return INFINITY;
            if (s == "-Infinity") return -INFINITY;
            char* end;
            double r = strtod(s.c_str(), &end);
            return *end ? NAN : r;
        }
        case T_OBJ: return to_number(to_primitive(v));
    }
    return NAN;
}
bool Interp::to_bool(const Value& v)
{
    switch (v.tag) {
        case T_BOOL: return v.b;
        case T_NUM: return !(v.n == 0 || std::isnan(v.n));
        case T_STR: return !v.s->s.empty();
        case T_OBJ: return true;
        default: return false;
    }
}
U16 Interp::num_to_string(double d, int radix)
{
    if (std::isnan(d)) return ascii("NaN");
    if (std::isinf(d)) return ascii(d > 0 ? "Infinity" : "-Infinity");
    if (d == 0) return ascii("0");
    if (radix != 10) {
        bool neg = d < 0;
        double t = std::floor(std::fabs(d));
        std::string s;
        if (t == 0) s = "0";
        while (t > 0) {
            int dig = (int)std::fmod(t, radix);
            s.insert(s.begin(), (char)(dig < 10 ? '0' + dig : 'a' + dig - 10));
            t = std::floor(t / radix);
        }
        if (neg) s.insert(s.begin(), '-');
        return ascii(s.c_str());
    }
    char buf[64];
    if (d == std::floor(d) && std::fabs(d) < 1e21) {
        snprintf(buf, sizeof buf, "%.0f", d);
        return ascii(buf);
    }
    for (int prec = 1; prec <= 17; ++prec) {  // shortest representation that round-trips
        snprintf(buf, sizeof buf, "%.*g", prec, d);
        if (strtod(buf, nullptr) == d) break;
    }
    return ascii(buf);
}
Value Interp::to_primitive(const Value& v)
{
    if (v.tag != T_OBJ) return v;
    Obj* o = v.o;
    if (o->kind == O_DATE) return Value(((DateObj*)o)->ms);
    if (o->kind == O_ARRAY) {
        ArrayObj* a = (ArrayObj*)o;
        U16 s;
        for (size_t i = 0; i < a->size(); ++i) {
            if (i) s += u',';
            const Value& e = a->at(i);
            if (e.tag != T_UNDEF && e.tag != T_NULL) s += to_string(e);
        }
        return Value::str(s);
    }
    if (o->kind == O_TYPED) {
        TypedObj* t = (TypedObj*)o;
        U16 s;
        for (size_t i = 0; i < t->len; ++i) {
            if (i) s += u',';
            s += to_string(typed_get(t, i));
        }
        return Value::str(s);
    }
    if (o->kind == O_FUNC) return Value::str(ascii("function () { [code] }"));
    Value ts = get_prop(v, A("toString"));
    if (ts.is_obj() && ts.o->kind == O_FUNC) {
        Value r = call(ts, v, nullptr, 0);
        if (r.tag != T_OBJ) return r;
    }
    if (o->proto == error_proto || (o->proto && o->proto->proto == error_proto)) {
        U16 s = to_string(get_prop(v, a_name));
        s += ascii(": ");
        s += to_string(get_prop(v, a_message));
        return Value::str(s);
    }
    return Value::str(ascii("[object Object]"));
}
U16 Interp::to_string(const Value& v)
{
    switch (v.tag) {
        case T_STR: return v.s->s;
        case T_NUM: return num_to_string(v.n);
        case T_UNDEF: return ascii("undefined");
        case T_NULL: return ascii("null");
        case T_BOOL: return ascii(v.b ? "true" : "false");
        case T_OBJ: return to_string(to_primitive(v));
    }
    return U16();
}
U16 Interp::type_of(const Value& v)
{
    switch (v.tag) {
        case T_UNDEF: return ascii("undefined");
        case T_NULL: return ascii("object");
        case T_BOOL: return ascii("boolean");
        case T_NUM: return ascii("number");
        case T_STR: return ascii("string");
        case T_OBJ: return ascii(v.o->kind == O_FUNC ? "function" : "object");
    }
    return U16();
}

// ---- typed arrays
Value Interp::typed_get(TypedObj* t, size_t i)
{
    const uint8_t* p = t->ptr();
    switch (t->ek) {
        case E_U8: return Value((double)p[i]);
        case E_I8: return Value((double)(int8_t)p[i]);
        case E_U16: {
            uint16_t v;
            memcpy(&v, p + 2 * i, 2);
            return Value((double)v);
        }
        case E_I16: {
            int16_t v;
            memcpy(&v, p + 2 * i, 2);
            return Value((double)v);
        }
        case E_U32: {
            uint32_t v;
            memcpy(&v, p + 4 * i, 4);
            return Value((double)v);
        }
        case E_I32: {
            int32_t v;
            memcpy(&v, p + 4 * i, 4);
            return Value((double)v);
        }
    }
    return Value();
}
void Interp::typed_set(TypedObj* t, size_t i, const Value& v)
{
    uint32_t u = to_uint32(v.tag == T_NUM ? v.n : to_number(v));  // modulo 2^32, then truncated to the element size
    uint8_t* p = t->ptr();
    switch (kElemSize[t->ek]) {
        case 1: p[i] = (uint8_t)u; break;
        case 2: {
            uint16_t w = (uint16_t)u;
            memcpy(p + 2 * i, &w, 2);
            break;
        }
        default: memcpy(p + 4 * i, &u, 4);
    }
}
TypedObj* Interp::new_typed(ElemKind ek, size_t len)
{
    BufferObj* b = new BufferObj();
    b->set_proto(buffer_proto);
    b->data.assign(len * kElemSize[ek], 0);
    return new_typed_on(ek, b, 0, len);
}
TypedObj* Interp::new_typed_on(ElemKind ek, BufferObj* b, size_t off, size_t len)
{
    TypedObj* t = new TypedObj();
    t->set_proto(typed_proto[ek]);
    t->buf = b;
    obj_retain(b);
    t->off = off;
    t->len = len;
    t->ek = ek;
    return t;
}
ArrayObj* Interp::new_array()
{
    ArrayObj* a = new ArrayObj();
    a->set_proto(array_proto);
    return a;
}
Obj* Interp::new_object()
{
    Obj* o = new Obj(O_PLAIN);
    o->set_proto(object_proto);
    return o;
}

// ---- property access
bool Interp::key_to_index(const Value& key, uint32_t& idx)
{
    if (key.tag == T_NUM) {
        double d = key.n;
        if (d >= 0 && d < 4294967295.0 && d == (double)(uint32_t)d) {
            idx = (uint32_t)d;
            return true;
        }
        return false;
    }
    if (key.tag == T_STR) {
        const U16& s = key.s->s;
        if (s.empty() || s.size() > 10) return false;
        if (s.size() > 1 && s[0] == u'0') return false;
        uint64_t v = 0;
        for (char16_t c : s) {
            if (c < u'0' || c > u'9') return false;
            v = v * 10 + (c - u'0');
        }
        if (v >= 4294967295ull) return false;
        idx = (uint32_t)v;
        return true;
    }
    return false;
}
Atom Interp::key_to_atom(const Value& key) { return g_atoms.get(narrow(to_string(key))); }

Value Interp::get_prop_obj(Obj* o, Atom a, const Value& self)
{
    for (Obj* p = o; p; p = p->proto) {
        if (Prop* pr = p->props.find(a)) {
            if (pr->getter) return call_function((FuncObj*)pr->getter, self, nullptr, 0);
            if (pr->setter) return Value();
            return pr->v;
        }
    }
    return Value();
}

Value Interp::get_prop(const Value& obj, Atom a)
{
    switch (obj.tag) {
        case T_OBJ: {
            Obj* o = obj.o;
            if (a == a_length) {
                if (o->kind == O_ARRAY) return Value((double)((ArrayObj*)o)->size());
                if (o->kind == O_TYPED) return Value((double)((TypedObj*)o)->len);
            }
            if (o->kind == O_TYPED) {
                TypedObj* t = (TypedObj*)o;
                if (a == a_buffer) return Value::obj(t->buf);
                if (a == a_byteOffset) return Value((double)t->off);
                if (a == a_byteLength) return Value((double)(t->len * kElemSize[t->ek]));
                if (a == a_BYTES_PER_ELEMENT) return Value((double)kElemSize[t->ek]);
            } else if (o->kind == O_BUFFER && a == a_byteLength) {
                return Value((double)((BufferObj*)o)->data.size());
            }
            return get_prop_obj(o, a, obj);
        }
        case T_STR:
            if (a == a_length) return Value((double)obj.s->s.size());
            return get_prop_obj(string_proto, a, obj);
        case T_NUM: return get_prop_obj(number_proto, a, obj);
        case T_BOOL: return get_prop_obj(object_proto, a, obj);
        default:
            throw_error("Cannot read properties of " + narrow(to_string(obj)) + " (reading '" + g_atoms.names[a] + "')",
                        "TypeError");
    }
}

void Interp::set_prop(const Value& obj, Atom a, const Value& v)
{
    if (obj.tag != T_OBJ) {
        if (obj.tag == T_UNDEF || obj.tag == T_NULL)
            throw_error("Cannot set properties of " + narrow(to_string(obj)) + " (setting '" + g_atoms.names[a] + "')",
                        "TypeError");
        return;  // primitives: silently ignored
    }
    Obj* o = obj.o;
    if (a == a_length) {
        if (o->kind == O_ARRAY) {
            ArrayObj* arr = (ArrayObj*)o;
            double d = to_number(v);
            if (!(d >= 0) || d != std::floor(d) || d > 4294967295.0) throw_error("Invalid array length", "RangeError");
            arr->el.resize(arr->head + (size_t)d);
            return;
        }
        if (o->kind == O_TYPED) return;
    }
    // own data property first (the common case), then accessors up the chain
    if (Prop* pr = o->props.find(a)) {
        if (pr->setter) {
            call_function((FuncObj*)pr->setter, obj, &v, 1);
            return;
        }
        if (pr->getter) return;
        pr->v = v;
        return;
    }
    for (Obj* p = o->proto; p; p = p->proto) {
        if (!p->has_accessors) continue;
        if (Prop* pr = p->props.find(a)) {
            if (pr->setter) {
                call_function((FuncObj*)pr->setter, obj, &v, 1);
                return;
            }
            if (pr->getter) return;
            break;
        }
    }
    o->props.insert(a)->v = v;
}

Value Interp::get_index(const Value& obj, const Value& key)
{
    uint32_t idx;
    if (obj.tag == T_OBJ) {
        Obj* o = obj.o;
        if (o->kind == O_TYPED) {
            TypedObj* t = (TypedObj*)o;
            if (key.tag == T_NUM) {
                double d = key.n;
                if (d >= 0 && d < (double)t->len && d == std::floor(d)) return typed_get(t, (size_t)d);
                if (d == std::floor(d) || std::isnan(d) || std::isinf(d)) return Value();  // out of range / not an index
            }
            if (key_to_index(key, idx)) return idx < t->len ? typed_get(t, idx) : Value();
        } else if (o->kind == O_ARRAY) {
            ArrayObj* a = (ArrayObj*)o;
            if (key_to_index(key, idx)) return idx < a->size() ? a->at(idx) : Value();
        } else if (key_to_index(key, idx)) {
            for (Obj* p = o; p; p = p->proto) {
                if (p->iprops) {
                    auto it = p->iprops->find(idx);
                    if (it != p->iprops->end()) return it->second;
                }
            }
            return Value();
        }
        return get_prop(obj, key_to_atom(key));
    }
    if (obj.tag == T_STR) {
        if (key_to_index(key, idx)) return idx < obj.s->s.size() ? Value::str(U16(1, obj.s->s[idx])) : Value();
        return get_prop(obj, key_to_atom(key));
    }
    if (obj.tag == T_UNDEF || obj.tag == T_NULL)
        throw_error("Cannot read properties of " + narrow(to_string(obj)) + " (reading '" + narrow(to_string(key)) + "')",
                    "TypeError");
    return get_prop(obj, key_to_atom(key));
}

void Interp::set_index(const Value& obj, const Value& key, const Value& v)
{
    uint32_t idx;
    if (obj.tag == T_OBJ) {
        Obj* o = obj.o;
        if (o->kind == O_TYPED) {
            TypedObj* t = (TypedObj*)o;
            if (key.tag == T_NUM) {
                double d = key.n;
                if (d >= 0 && d < (double)t->len && d == std::floor(d)) {
                    typed_set(t, (size_t)d, v);
                    return;
                }
                // canonical numeric keys never become ordinary properties of a typed array: the write is dropped
                // (negative, fractional, NaN and past-the-end indices alike)
                return;
            }
            if (key_to_index(key, idx)) {
                if (idx < t->len) typed_set(t, idx, v);
                return;
            }
        } else if (o->kind == O_ARRAY) {
            ArrayObj* a = (ArrayObj*)o;
            if (key_to_index(key, idx)) {
                if (idx >= a->size()) {
                    if (idx > (1u << 28)) throw_error("array index too large for this interpreter", "RangeError");
                    a->el.resize(a->head + idx + 1);
                }
                a->at(idx) = v;
                return;
            }
        } else if (key_to_index(key, idx)) {
            if (!o->iprops) o->iprops = new std::unordered_map<uint32_t, Value>();
            (*o->iprops)[idx] = v;
            return;
        }
        set_prop(obj, key_to_atom(key), v);
        return;
    }
    if (obj.tag == T_UNDEF || obj.tag == T_NULL)
        throw_error("Cannot set properties of " + narrow(to_string(obj)) + " (setting '" + narrow(to_string(key)) + "')",
                    "TypeError");
}

bool Interp::has_property(Obj* o, const Value& key)
{
    uint32_t idx;
    if (key_to_index(key, idx)) {
        if (o->kind == O_TYPED) return idx < ((TypedObj*)o)->len;
        if (o->kind == O_ARRAY) return idx < ((ArrayObj*)o)->size();
        for (Obj* p = o; p; p = p->proto)
            if (p->iprops && p->iprops->count(idx)) return true;
        return false;
    }
    Atom a = key_to_atom(key);
    if (a == a_length && (o->kind == O_ARRAY || o->kind == O_TYPED)) return true;
    for (Obj* p = o; p; p = p->proto)
        if (p->props.find(a)) return true;
    return false;
}

// ---- operators
bool Interp::strict_equals(const Value& a, const Value& b)
{
    if (a.tag != b.tag) return false;
    switch (a.tag) {
        case T_UNDEF:
        case T_NULL: return true;
        case T_BOOL: return a.b == b.b;
        case T_NUM: return a.n == b.n;
        case T_STR: return a.s == b.s || a.s->s == b.s->s;
        case T_OBJ: return a.o == b.o;
    }
    return false;
}
bool Interp::loose_equals(const Value& a, const Value& b)
{
    if (a.tag == b.tag) return strict_equals(a, b);
    bool an = a.tag == T_UNDEF || a.tag == T_NULL, bn = b.tag == T_UNDEF || b.tag == T_NULL;
    if (an || bn) return an && bn;
    if (a.tag == T_OBJ && b.tag != T_OBJ) return loose_equals(to_primitive(a), b);
    if (b.tag == T_OBJ && a.tag != T_OBJ) return loose_equals(a, to_primitive(b));
    return to_number(a) == to_number(b);
}
bool Interp::instance_of(const Value& v, const Value& ctor)
{
    if (!ctor.is_obj() || ctor.o->kind != O_FUNC) throw_error("Right-hand side of 'instanceof' is not callable", "TypeError");
    if (!v.is_obj()) return false;
    Value pv = get_prop(ctor, a_prototype);
    if (!pv.is_obj()) return false;
    for (Obj* p = v.o->proto; p; p = p->proto)
        if (p == pv.o) return true;
    return false;
}

Value Interp::binary_op(Op op, const Value& a, const Value& b)
{
    if (a.tag == T_NUM && b.tag == T_NUM) {
        double x = a.n, y = b.n;
        switch (op) {
            case OP_ADD: return Value(x + y);
            case OP_SUB: return Value(x - y);
            case OP_MUL: return Value(x * y);
            case OP_DIV: return Value(x / y);
            case OP_MOD: return Value(std::fmod(x, y));
            case OP_POW: return Value(std::pow(x, y));
            case OP_SHL: return Value((double)(int32_t)((uint32_t)to_int32(x) << (to_uint32(y) & 31)));
            case OP_SHR: return Value((double)(to_int32(x) >> (to_uint32(y) & 31)));
            case OP_USHR: return Value((double)(to_uint32(x) >> (to_uint32(y) & 31)));
            case OP_AND: return Value((double)(to_int32(x) & to_int32(y)));
            case OP_OR: return Value((double)(to_int32(x) | to_int32(y)));
            case OP_XOR: return Value((double)(to_int32(x) ^ to_int32(y)));
            case OP_LT: return Value::boolean(x < y);
            case OP_GT: return Value::boolean(x > y);
            case OP_LE: return Value::boolean(x <= y);
            case OP_GE: return Value::boolean(x >= y);
            case OP_EQ:
            case OP_SEQ: return Value::boolean(x == y);
            case OP_NE:
            case OP_SNE: return Value::boolean(x != y);
            default: break;
        }
    }
    switch (op) {
        case OP_ADD: {
            Value pa = to_primitive(a), pb = to_primitive(b);
            if (pa.tag == T_STR || pb.tag == T_STR) {
                U16 s = to_string(pa);
                s += to_string(pb);
                return Value::str(s);
            }
            return Value(to_number(pa) + to_number(pb));
        }
        case OP_SUB: return Value(to_number(a) - to_number(b));
        case OP_MUL: return Value(to_number(a) * to_number(b));
        case OP_DIV: return Value(to_number(a) / to_number(b));
        case OP_MOD: return Value(std::fmod(to_number(a), to_number(b)));
        case OP_POW: return Value(std::pow(to_number(a), to_number(b)));
        case OP_SHL: return Value((double)(int32_t)((uint32_t)to_int32(a) << (to_uint32(to_number(b)) & 31)));
        case OP_SHR: return Value((double)(to_int32(a) >> (to_uint32(to_number(b)) & 31)));
        case OP_USHR: return Value((double)(to_uint32(to_number(a)) >> (to_uint32(to_number(b)) & 31)));
        case OP_AND: return Value((double)(to_int32(a) & to_int32(b)));
        case OP_OR: return Value((double)(to_int32(a) | to_int32(b)));
        case OP_XOR: return Value((double)(to_int32(a) ^ to_int32(b)));
        case OP_LT:
        case OP_GT:
        case OP_LE:
        case OP_GE: {
            Value pa = to_primitive(a), pb = to_primitive(b);
            if (pa.tag == T_STR && pb.tag == T_STR) {
                int c = pa.s->s.compare(pb.s->s);
                return Value::boolean(op == OP_LT ? c < 0 : op == OP_GT ? c > 0 : op == OP_LE ? c <= 0 : c >= 0);
            }
            double x = to_number(pa), y = to_number(pb);
            return Value::boolean(op == OP_LT ? x < y : op == OP_GT ? x > y : op == OP_LE ? x <= y : x >= y);
        }
        case OP_EQ: return Value::boolean(loose_equals(a, b));
        case OP_NE: return Value::boolean(!loose_equals(a, b));
        case OP_SEQ: return Value::boolean(strict_equals(a, b));
        case OP_SNE: return Value::boolean(!strict_equals(a, b));
        case OP_INSTANCEOF: return Value::boolean(instance_of(a, b));
        case OP_IN:
            if (!b.is_obj()) throw_error("Cannot use 'in' operator to search in a primitive", "TypeError");
            return Value::boolean(has_property(b.o, a));
        default: break;
    }
    throw_error("unsupported operator");
}

// ---- variables
Env* Interp::this_env(Env* env) { return env; }

Value& Interp::var_ref(Node* id, Env* env, bool& is_global)
{
    is_global = false;
    Env* e = env;
    for (int h = id->hops; h > 0; --h) e = e->parent;
    return e->slots[id->slot];
}

void Interp::assign_to(Node* target, const Value& v, Env* env)
{
    switch (target->kind) {
        case N_IDENT:
            if (target->slot >= 0) {
                bool g;
                var_ref(target, env, g) = v;
            } else {
                global->props.insert(target->atom)->v = v;
            }
            return;
        case N_MEMBER: {
            Value o = eval(target->a, env);
            set_prop(o, target->atom, v);
            return;
        }
        case N_INDEX: {
            Value o = eval(target->a, env);
            Value k = eval(target->b, env);
            set_index(o, k, v);
            return;
        }
        default: throw_error("invalid assignment target", "SyntaxError");
    }
}

// ---- functions
Value Interp::make_closure(FuncNode* fn, Env* env)
{
    FuncObj* f = new FuncObj();
    f->set_proto(function_proto);
    f->node = fn;
    f->env = env;
    env_retain(env);
    f->name = fn->name;
    return Value::obj(f);
}

Value Interp::call(const Value& fn, const Value& self, const Value* args, int argc)
{
    if (!fn.is_obj() || fn.o->kind != O_FUNC) throw_error(narrow(to_string(fn)) + " is not a function", "TypeError");
    return call_function((FuncObj*)fn.o, self, args, argc);
}

Value Interp::call_function(FuncObj* f, const Value& self, const Value* args, int argc)
{
    if (f->native) return f->native(*this, self, args, argc);
    if (f->is_class && self.tag != T_OBJ) throw_error("Class constructor cannot be invoked without 'new'", "TypeError");
    FuncNode* fn = f->node;
    if (++depth > 2000) {
        depth = 0;
        throw_error("Maximum call stack size exceeded", "RangeError");
    }
    Env* env = new Env();
    env->rc = 1;
    env->parent = f->env;
    env_retain(f->env);
    env->slots.resize(fn->locals.size());
    if (fn->is_arrow)
        env->this_val = f->env ? f->env->this_val : Value();
    else
        env->this_val = self;
    int np = fn->n_params;
    for (int i = 0; i < np; ++i) {
        if (i < argc && !args[i].is_undef())
            env->slots[i] = args[i];
        else if (fn->param_defaults[i])
            env->slots[i] = eval(fn->param_defaults[i], env);
    }
    if (fn->has_self) {
        int sl = fn->slot_of(fn->self_name);
        if (sl >= np) env->slots[sl] = Value::obj(f);
    }
    for (Node* d : fn->func_decls) env->slots[d->a->slot] = make_closure(d->fn, env);
    Value result;
    try {
        if (fn->expr_body) {
            result = eval(fn->body, env);
        } else {
            Completion c = exec(fn->body, env);
            if (c == C_RETURN) {
                result = std::move(ret);
                ret = Value();
            }
        }
    } catch (...) {
        --depth;
        env_release(env);
        throw;
    }
    --depth;
    env_release(env);
    return result;
}

Value Interp::construct(const Value& fn, const Value* args, int argc)
{
    if (!fn.is_obj() || fn.o->kind != O_FUNC) throw_error(narrow(to_string(fn)) + " is not a constructor", "TypeError");
    FuncObj* f = (FuncObj*)fn.o;
    if (f->native) {
        Value marker = Value::obj(f);  // natives get the constructor itself as `this` when called through new
        return f->native(*this, marker, args, argc);
    }
    Obj* o = new Obj(O_PLAIN);
    Value self = Value::obj(o);
    Value pv = get_prop(fn, a_prototype);
    o->set_proto(pv.is_obj() ? pv.o : object_proto);
    // instance fields
    Value fields = get_prop(fn, A("__fields"));
    if (fields.is_obj() && fields.o->kind == O_ARRAY) {
        ArrayObj* fa = (ArrayObj*)fields.o;
        for (size_t i = 0; i + 1 < fa->size(); i += 2) {
            Value init = fa->at(i + 1);
            Value v = init.is_obj() ? call(init, self, nullptr, 0) : Value();
            set_prop(self, key_to_atom(fa->at(i)), v);
        }
    }
    Value r = call_function(f, self, args, argc);
    if (r.is_obj()) return r;
    return self;
}

Value Interp::eval_class(Node* n, Env* env)
{
    // constructor
    FuncNode* ctor_node = nullptr;
    Atom a_ctor = a_constructor;
    for (auto& pd : n->props)
        if (!pd.is_static && pd.accessor == 0 && pd.key == a_ctor) ctor_node = pd.value->fn;
    FuncObj* cls = new FuncObj();
    Value clsv = Value::obj(cls);
    cls->set_proto(function_proto);
    cls->is_class = true;
    cls->env = env;
    env_retain(env);
    if (ctor_node) {
        cls->node = ctor_node;
    } else {
        static FuncNode* empty = nullptr;
        if (!empty) {
            empty = new FuncNode();
            empty->name = "<default constructor>";
            empty->body = new Node(N_BLOCK);
        }
        cls->node = empty;
    }
    cls->name = n->atom ? g_atoms.names[n->atom] : "";
    Obj* proto = new_object();
    Value protov = Value::obj(proto);
    cls->props.insert(a_prototype)->v = protov;
    proto->props.insert(a_constructor)->v = clsv;
    if (n->a) assign_to(n->a, clsv, env);  // the inner name
    ArrayObj* fields = nullptr;
    for (auto& pd : n->props) {
        if (!pd.is_static && pd.accessor == 0 && pd.key == a_ctor) continue;
        Obj* target = pd.is_static ? (Obj*)cls : proto;
        if (pd.accessor == 3) {  // field
            Value init = pd.value ? make_closure(pd.value->fn, env) : Value();
            if (pd.is_static) {
                Value v = init.is_obj() ? call(init, clsv, nullptr, 0) : Value();
                cls->props.insert(pd.key)->v = v;
            } else {
                if (!fields) {
                    fields = new_array();
                    cls->props.insert(A("__fields"))->v = Value::obj(fields);
                }
                fields->el.push_back(Value::str(from_utf8(g_atoms.names[pd.key])));
                fields->el.push_back(init);
            }
            continue;
        }
        Value fv = make_closure(pd.value->fn, env);
        Prop* pr = target->props.insert(pd.key);
        if (pd.accessor == 1) {
            pr->getter = fv.o;
            obj_retain(fv.o);
            target->has_accessors = true;
        } else if (pd.accessor == 2) {
            pr->setter = fv.o;
            obj_retain(fv.o);
            target->has_accessors = true;
        } else
            pr->v = fv;
    }
    return clsv;
}

// ---- evaluation
Value Interp::eval(Node* n, Env* env)
{
    switch (n->kind) {
        case N_NUM: return Value(n->num);
        case N_STR: return n->strv;
        case N_IDENT: {
            if (n->slot >= 0) {
                Env* e = env;
                for (int h = n->hops; h > 0; --h) e = e->parent;
                return e->slots[n->slot];
            }
            if (Prop* p = global->props.find(n->atom)) return p->v;
            throw_error(g_atoms.names[n->atom] + " is not defined", "ReferenceError");
        }
        case N_THIS: return env->this_val;
        case N_NULL: return Value::null();
        case N_UNDEF: return Value();
        case N_TRUE: return Value::boolean(true);
        case N_FALSE: return Value::boolean(false);
        case N_ARRAY: {
            ArrayObj* a = new_array();
            Value av = Value::obj(a);
            a->el.reserve(n->list.size());
            for (Node* e : n->list) a->el.push_back(eval(e, env));
            return av;
        }
        case N_OBJECT: {
            Obj* o = new_object();
            Value ov = Value::obj(o);
            for (auto& pd : n->props) {
                Value v = pd.value->kind == N_FUNC ? make_closure(pd.value->fn, env) : eval(pd.value, env);
                if (pd.computed) {
                    set_index(ov, eval(pd.computed, env), v);
                } else if (pd.is_index) {
                    set_index(ov, Value((double)pd.index), v);
                } else if (pd.accessor) {
                    Prop* pr = o->props.insert(pd.key);
                    obj_retain(v.o);
                    if (pd.accessor == 1) pr->getter = v.o;
                    else pr->setter = v.o;
                    o->has_accessors = true;
                } else
                    o->props.insert(pd.key)->v = v;
            }
            return ov;
        }
        case N_FUNC: return make_closure(n->fn, env);
        case N_CLASS: return eval_class(n, env);
        case N_MEMBER: {
            Value o = eval(n->a, env);
            return get_prop(o, n->atom);
        }
        case N_INDEX: {
            Value o = eval(n->a, env);
            Value k = eval(n->b, env);
            return get_index(o, k);
        }
        case N_CALL: {
            Value self, fn;
            Node* cal = n->a;
            if (cal->kind == N_MEMBER) {
                self = eval(cal->a, env);
                fn = get_prop(self, cal->atom);
                if (!fn.is_obj() || fn.o->kind != O_FUNC) {
                    throw_error(g_atoms.names[cal->atom] + " is not a function (line " + std::to_string(n->line) + ")", "TypeError");
                }
            } else if (cal->kind == N_INDEX) {
                self = eval(cal->a, env);
                fn = get_index(self, eval(cal->b, env));
            } else
                fn = eval(cal, env);
            size_t argc = n->list.size();
            Value stack_args[6];
            std::vector<Value> heap_args;
            Value* args = stack_args;
            if (argc > 6) {
                heap_args.resize(argc);
                args = heap_args.data();
            }
            for (size_t i = 0; i < argc; ++i) args[i] = eval(n->list[i], env);
            return call(fn, self, args, (int)argc);
        }
        case N_NEW: {
            Value fn = eval(n->a, env);
            size_t argc = n->list.size();
            std::vector<Value> args(argc);
            for (size_t i = 0; i < argc; ++i) args[i] = eval(n->list[i], env);
            return construct(fn, args.data(), (int)argc);
        }
        case N_UNARY: {
            if (n->op == OP_TYPEOF) {
                if (n->a->kind == N_IDENT && n->a->slot < 0 && !global->props.find(n->a->atom)) return Value::str(ascii("undefined"));
                return Value::str(type_of(eval(n->a, env)));
            }
            if (n->op == OP_DELETE) {
                if (n->a->kind == N_MEMBER) {
                    Value o = eval(n->a->a, env);
                    if (o.is_obj()) o.o->props.erase(n->a->atom);
                } else if (n->a->kind == N_INDEX) {
                    Value o = eval(n->a->a, env);
                    Value k = eval(n->a->b, env);
                    uint32_t idx;
                    if (o.is_obj()) {
                        if (key_to_index(k, idx)) {
                            if (o.o->kind == O_ARRAY) {
                                if (idx < ((ArrayObj*)o.o)->size()) ((ArrayObj*)o.o)->at(idx) = Value();
                            } else if (o.o->iprops)
                                o.o->iprops->erase(idx);
                        } else
                            o.o->props.erase(key_to_atom(k));
                    }
                }
                return Value::boolean(true);
            }
            Value v = eval(n->a, env);
            switch (n->op) {
                case OP_NOT: return Value::boolean(!to_bool(v));
                case OP_BITNOT: return Value((double)(~to_int32(v)));
                case OP_NEG: return Value(-to_number(v));
                case OP_PLUS: return Value(to_number(v));
                case OP_VOID: return Value();
                default: break;
            }
            throw_error("unsupported unary operator");
        }
        case N_UPDATE: {
            Node* t = n->a;
            double delta = n->op == OP_ADD ? 1 : -1;
            if (t->kind == N_IDENT && t->slot >= 0) {
                bool g;
                Value& ref = var_ref(t, env, g);
                double old = ref.tag == T_NUM ? ref.n : to_number(ref);
                ref = Value(old + delta);
                return Value(n->prefix ? old + delta : old);
            }
            if (t->kind == N_MEMBER) {
                Value o = eval(t->a, env);
                double old = to_number(get_prop(o, t->atom));
                set_prop(o, t->atom, Value(old + delta));
                return Value(n->prefix ? old + delta : old);
            }
            if (t->kind == N_INDEX) {
                Value o = eval(t->a, env);
                Value k = eval(t->b, env);
                double old = to_number(get_index(o, k));
                set_index(o, k, Value(old + delta));
                return Value(n->prefix ? old + delta : old);
            }
            double old = to_number(eval(t, env));
            assign_to(t, Value(old + delta), env);
            return Value(n->prefix ? old + delta : old);
        }
        case N_BINARY: {
            Value a = eval(n->a, env);
            Value b = eval(n->b, env);
            return binary_op(n->op, a, b);
        }
        case N_LOGICAL: {
            Value a = eval(n->a, env);
            if (n->op == OP_LAND) return to_bool(a) ? eval(n->b, env) : a;
            if (n->op == OP_LOR) return to_bool(a) ? a : eval(n->b, env);
            return (a.tag == T_UNDEF || a.tag == T_NULL) ? eval(n->b, env) : a;
        }
        case N_ASSIGN: {
            Node* t = n->a;
            if (n->op == OP_NONE) {
                if (t->kind == N_IDENT && t->slot >= 0) {
                    Value v = eval(n->b, env);
                    bool g;
                    var_ref(t, env, g) = v;
                    return v;
                }
                if (t->kind == N_MEMBER) {
                    Value o = eval(t->a, env);
                    Value v = eval(n->b, env);
                    set_prop(o, t->atom, v);
                    return v;
                }
                if (t->kind == N_INDEX) {
                    Value o = eval(t->a, env);
                    Value k = eval(t->b, env);
                    Value v = eval(n->b, env);
                    set_index(o, k, v);
                    return v;
                }
                Value v = eval(n->b, env);
                assign_to(t, v, env);
                return v;
            }
            // compound: the target's object and key are evaluated once
            Value o, k, cur;
            if (t->kind == N_MEMBER) {
                o = eval(t->a, env);
                cur = get_prop(o, t->atom);
            } else if (t->kind == N_INDEX) {
                o = eval(t->a, env);
                k = eval(t->b, env);
                cur = get_index(o, k);
            } else
                cur = eval(t, env);
            Value v;
            if (n->op == OP_LAND) {
                if (!to_bool(cur)) return cur;
                v = eval(n->b, env);
            } else if (n->op == OP_LOR) {
                if (to_bool(cur)) return cur;
                v = eval(n->b, env);
            } else if (n->op == OP_NULLISH) {
                if (!(cur.tag == T_UNDEF || cur.tag == T_NULL)) return cur;
                v = eval(n->b, env);
            } else {
                Value rhs = eval(n->b, env);
                v = binary_op(n->op, cur, rhs);
            }
            if (t->kind == N_MEMBER) set_prop(o, t->atom, v);
            else if (t->kind == N_INDEX) set_index(o, k, v);
            else assign_to(t, v, env);
            return v;
        }
        case N_COND: return to_bool(eval(n->a, env)) ? eval(n->b, env) : eval(n->c, env);
        case N_SEQ: {
            Value v;
            for (Node* e : n->list) v = eval(e, env);
            return v;
        }
        default: break;
    }
    throw_error("cannot evaluate node");
}

Completion Interp::exec_list(const std::vector<Node*>& l, Env* env)
{
    for (Node* s : l) {
        Completion c = exec(s, env);
        if (c != C_NORMAL) return c;
    }
    return C_NORMAL;
}

Completion Interp::exec(Node* n, Env* env)
{
    switch (n->kind) {
        case N_EXPR: eval(n->a, env); return C_NORMAL;
        case N_VAR:
            for (size_t i = 0; i + 1 < n->list.size(); i += 2)
                if (n->list[i + 1]) assign_to(n->list[i], eval(n->list[i + 1], env), env);
            return C_NORMAL;
        case N_FUNCDECL:
        case N_EMPTY: return C_NORMAL;
        case N_BLOCK: return exec_list(n->list, env);
        case N_IF:
            if (to_bool(eval(n->a, env))) return exec(n->b, env);
            if (n->c) return exec(n->c, env);
            return C_NORMAL;
        case N_FOR: {
            if (n->a) exec(n->a, env);
            for (;;) {
                if (n->b && !to_bool(eval(n->b, env))) break;
                Completion c = exec(n->d, env);
                if (c == C_BREAK) break;
                if (c == C_RETURN) return c;
                if (n->c) eval(n->c, env);
            }
            return C_NORMAL;
        }
        case N_WHILE:
            while (to_bool(eval(n->a, env))) {
                Completion c = exec(n->b, env);
                if (c == C_BREAK) break;
                if (c == C_RETURN) return c;
            }
            return C_NORMAL;
        case N_DOWHILE:
            do {
                Completion c = exec(n->b, env);
                if (c == C_BREAK) break;
                if (c == C_RETURN) return c;
            } while (to_bool(eval(n->a, env)));
            return C_NORMAL;
        case N_FORIN: {
            Value o = eval(n->b, env);
            std::vector<Value> items;
            if (o.is_obj()) {
                Obj* ob = o.o;
                if (n->for_of) {
                    if (ob->kind == O_ARRAY)
                        for (size_t i = 0; i < ((ArrayObj*)ob)->size(); ++i) items.push_back(((ArrayObj*)ob)->at(i));
                    else if (ob->kind == O_TYPED)
                        for (size_t i = 0; i < ((TypedObj*)ob)->len; ++i) items.push_back(typed_get((TypedObj*)ob, i));
                    else
                        throw_error("object is not iterable", "TypeError");
                } else {
                    if (ob->kind == O_ARRAY)
                        for (size_t i = 0; i < ((ArrayObj*)ob)->size(); ++i) items.push_back(Value::str(num_to_string((double)i)));
                    else if (ob->kind == O_TYPED)
                        for (size_t i = 0; i < ((TypedObj*)ob)->len; ++i) items.push_back(Value::str(num_to_string((double)i)));
                    if (ob->iprops) {
                        std::vector<uint32_t> ks;
                        for (auto& kv : *ob->iprops) ks.push_back(kv.first);
                        std::sort(ks.begin(), ks.end());
                        for (uint32_t k : ks) items.push_back(Value::str(num_to_string((double)k)));
                    }
                    for (Atom k : ob->props.keys) items.push_back(Value::str(from_utf8(g_atoms.names[k])));
                }
            } else if (o.is_str() && n->for_of) {
                for (char16_t c : o.s->s) items.push_back(Value::str(U16(1, c)));
            }
            for (auto& it : items) {
                assign_to(n->a, it, env);
                Completion c = exec(n->c, env);
                if (c == C_BREAK) break;
                if (c == C_RETURN) return c;
            }
            return C_NORMAL;
        }
        case N_RETURN:
            ret = n->a ? eval(n->a, env) : Value();
            return C_RETURN;
        case N_BREAK: return C_BREAK;
        case N_CONTINUE: return C_CONTINUE;
        case N_THROW: throw JsThrow{eval(n->a, env)};
        case N_SWITCH: {
            Value v = eval(n->a, env);
            int start = -1;
            for (size_t i = 0; i < n->cases.size(); ++i)
                if (n->cases[i].test && strict_equals(v, eval(n->cases[i].test, env))) {
                    start = (int)i;
                    break;
                }
            if (start < 0)
                for (size_t i = 0; i < n->cases.size(); ++i)
                    if (!n->cases[i].test) start = (int)i;
            if (start < 0) return C_NORMAL;
            for (size_t i = (size_t)start; i < n->cases.size(); ++i) {
                Completion c = exec_list(n->cases[i].body, env);
                if (c == C_BREAK) return C_NORMAL;
                if (c != C_NORMAL) return c;  // continue belongs to the enclosing loop
            }
            return C_NORMAL;
        }
        case N_TRY: {
            Completion c = C_NORMAL;
            bool rethrow = false;
            JsThrow pending;
            try {
                c = exec(n->a, env);
            } catch (JsThrow& t) {
                int saved_depth = depth;
                (void)saved_depth;
                if (n->b) {
                    if (n->d) assign_to(n->d, t.v, env);
                    try {
                        c = exec(n->b, env);
                    } catch (JsThrow& t2) {
                        rethrow = true;
                        pending = t2;
                    }
                } else {
                    rethrow = true;
                    pending = t;
                }
            }
            if (n->c) {
                Value saved_ret = ret;
                Completion fc = exec(n->c, env);
                if (fc != C_NORMAL) return fc;
                ret = saved_ret;
            }
            if (rethrow) throw pending;
            return c;
        }
        default: eval(n, env); return C_NORMAL;
    }
}

// =====================================================================================================
// Built-ins
// =====================================================================================================
#define ARG(i) ((i) < argc ? args[i] : Value())

FuncObj* Interp::make_native(const char* name, NativeFn fn, int ctor_kind)
{
    FuncObj* f = new FuncObj();
    f->set_proto(function_proto);
    f->native = fn;
    f->ctor_kind = ctor_kind;
    f->name = name;
    return f;
}
void Interp::def_native(Obj* target, const char* name, NativeFn fn, int ctor_kind)
{
    target->props.insert(A(name))->v = Value::obj(make_native(name, fn, ctor_kind));
}

static double arg_int(Interp& I, const Value& v, double dflt)
{
    if (v.is_undef()) return dflt;
    double d = I.to_number(v);
    if (std::isnan(d)) return 0;
    return std::trunc(d);
}
static size_t rel_index(double v, size_t len)
{
    if (v < 0) {
        v += (double)len;
        if (v < 0) v = 0;
    }
    if (v > (double)len) v = (double)len;
    return (size_t)v;
}

// ---- typed array constructor: (length) | (array-like) | (buffer, byteOffset?, length?)
static Value typed_ctor(Interp& I, const Value& self, const Value* args, int argc)
{
    if (!self.is_obj() || self.o->kind != O_FUNC) I.throw_error("Constructor requires 'new'", "TypeError");
    ElemKind ek = (ElemKind)((FuncObj*)self.o)->ctor_kind;
    Value a0 = ARG(0);
    if (a0.is_obj()) {
        Obj* o = a0.o;
        if (o->kind == O_BUFFER) {
            BufferObj* b = (BufferObj*)o;
            double off = arg_int(I, ARG(1), 0);
            int es = kElemSize[ek];
            if (off < 0 || off > (double)b->data.size() || std::fmod(off, es) != 0)
                I.throw_error("start offset of typed array should be a multiple of its element size and inside the buffer", "RangeError");
            size_t len;
            if (ARG(2).is_undef()) {
                if ((b->data.size() - (size_t)off) % es) I.throw_error("byte length of typed array should be a multiple of its element size", "RangeError");
                len = (b->data.size() - (size_t)off) / es;
            } else {
                double l = arg_int(I, ARG(2), 0);
                if (l < 0 || off + l * es > (double)b->data.size()) I.throw_error("Invalid typed array length: " + narrow(I.to_string(ARG(2))), "RangeError");
                len = (size_t)l;
            }
            return Value::obj(I.new_typed_on(ek, b, (size_t)off, len));
        }
        if (o->kind == O_TYPED) {
            TypedObj* s = (TypedObj*)o;
            TypedObj* t = I.new_typed(ek, s->len);
            Value tv = Value::obj(t);
            if (s->ek == ek)
                memcpy(t->ptr(), s->ptr(), s->len * kElemSize[ek]);
            else
                for (size_t i = 0; i < s->len; ++i) I.typed_set(t, i, Interp::typed_get(s, i));
            return tv;
        }
        if (o->kind == O_ARRAY) {
            ArrayObj* s = (ArrayObj*)o;
            TypedObj* t = I.new_typed(ek, s->size());
            Value tv = Value::obj(t);
            for (size_t i = 0; i < s->size(); ++i) I.typed_set(t, i, s->at(i));
            return tv;
        }
        // array-like object with a length
        double l = I.to_number(I.get_prop(a0, I.a_length));
        size_t len = std::isnan(l) || l < 0 ? 0 : (size_t)l;
        TypedObj* t = I.new_typed(ek, len);
        Value tv = Value::obj(t);
        for (size_t i = 0; i < len; ++i) I.typed_set(t, i, I.get_index(a0, Value((double)i)));
        return tv;
    }
    double l = a0.is_undef() ? 0 : I.to_number(a0);
    if (std::isnan(l)) l = 0;
    if (l < 0 || l != std::floor(l) || l > 4294967296.0) I.throw_error("Invalid typed array length: " + narrow(I.to_string(a0)), "RangeError");
    return Value::obj(I.new_typed(ek, (size_t)l));
}
static TypedObj* this_typed(Interp& I, const Value& self)
{
    if (!self.is_obj() || self.o->kind != O_TYPED) I.throw_error("this is not a typed array", "TypeError");
    return (TypedObj*)self.o;
}
static Value typed_set_m(Interp& I, const Value& self, const Value* args, int argc)
{
    TypedObj* t = this_typed(I, self);
    double off = arg_int(I, ARG(1), 0);
    Value src = ARG(0);
    if (!src.is_obj()) return Value();
    size_t n;
    if (src.o->kind == O_TYPED) n = ((TypedObj*)src.o)->len;
    else if (src.o->kind == O_ARRAY) n = ((ArrayObj*)src.o)->size();
    else {
        double l = I.to_number(I.get_prop(src, I.a_length));
        n = std::isnan(l) || l < 0 ? 0 : (size_t)l;
    }
    if (off < 0 || off + (double)n > (double)t->len) I.throw_error("offset is out of bounds", "RangeError");
    size_t o = (size_t)off;
    if (src.o->kind == O_TYPED) {
        TypedObj* s = (TypedObj*)src.o;
        if (s->ek == t->ek) {
            memmove(t->ptr() + o * kElemSize[t->ek], s->ptr(), n * kElemSize[t->ek]);
        } else {
            std::vector<Value> tmp(n);
            for (size_t i = 0; i < n; ++i) tmp[i] = Interp::typed_get(s, i);
            for (size_t i = 0; i < n; ++i) I.typed_set(t, o + i, tmp[i]);
        }
    } else if (src.o->kind == O_ARRAY) {
        ArrayObj* s = (ArrayObj*)src.o;
        for (size_t i = 0; i < n; ++i) I.typed_set(t, o + i, s->at(i));
    } else {
        for (size_t i = 0; i < n; ++i) I.typed_set(t, o + i, I.get_index(src, Value((double)i)));
    }
    return Value();
}
static Value typed_subarray(Interp& I, const Value& self, const Value* args, int argc)
{
    TypedObj* t = this_typed(I, self);
    size_t b = rel_index(arg_int(I, ARG(0), 0), t->len);
    size_t e = rel_index(arg_int(I, ARG(1), (double)t->len), t->len);
    if (e < b) e = b;
    return Value::obj(I.new_typed_on(t->ek, t->buf, t->off + b * kElemSize[t->ek], e - b));
}
static Value typed_slice(Interp& I, const Value& self, const Value* args, int argc)
{
    TypedObj* t = this_typed(I, self);
    size_t b = rel_index(arg_int(I, ARG(0), 0), t->len);
    size_t e = rel_index(arg_int(I, ARG(1), (double)t->len), t->len);
    if (e < b) e = b;
    TypedObj* r = I.new_typed(t->ek, e - b);
    memcpy(r->ptr(), t->ptr() + b * kElemSize[t->ek], (e - b) * kElemSize[t->ek]);
    return Value::obj(r);
}
static Value typed_fill(Interp& I, const Value& self, const Value* args, int argc)
{
    TypedObj* t = this_typed(I, self);
    size_t b = rel_index(arg_int(I, ARG(1), 0), t->len);
    size_t e = rel_index(arg_int(I, ARG(2), (double)t->len), t->len);
    for (size_t i = b; i < e; ++i) I.typed_set(t, i, ARG(0));
    return self;
}
static Value typed_map(Interp& I, const Value& self, const Value* args, int argc)
{
    TypedObj* t = this_typed(I, self);
    TypedObj* r = I.new_typed(t->ek, t->len);
    Value rv = Value::obj(r);
    for (size_t i = 0; i < t->len; ++i) {
        Value cb_args[3] = {Interp::typed_get(t, i), Value((double)i), self};
        I.typed_set(r, i, I.call(ARG(0), Value(), cb_args, 3));
    }
    return rv;
}
static Value typed_forEach(Interp& I, const Value& self, const Value* args, int argc)
{
    TypedObj* t = this_typed(I, self);
    for (size_t i = 0; i < t->len; ++i) {
        Value cb_args[3] = {Interp::typed_get(t, i), Value((double)i), self};
        I.call(ARG(0), Value(), cb_args, 3);
    }
    return Value();
}
static Value typed_join(Interp& I, const Value& self, const Value* args, int argc)
{
    TypedObj* t = this_typed(I, self);
    U16 sep = ARG(0).is_undef() ? ascii(",") : I.to_string(ARG(0));
    U16 s;
    for (size_t i = 0; i < t->len; ++i) {
        if (i) s += sep;
        s += I.to_string(Interp::typed_get(t, i));
    }
    return Value::str(s);
}
static Value typed_indexOf(Interp& I, const Value& self, const Value* args, int argc)
{
    TypedObj* t = this_typed(I, self);
    for (size_t i = 0; i < t->len; ++i)
        if (I.strict_equals(Interp::typed_get(t, i), ARG(0))) return Value((double)i);
    return Value(-1.0);
}
static Value buffer_ctor(Interp& I, const Value&, const Value* args, int argc)
{
    BufferObj* b = new BufferObj();
    b->set_proto(I.buffer_proto);
    double l = arg_int(I, ARG(0), 0);
    if (l < 0 || l > 4294967296.0) I.throw_error("Array buffer allocation failed", "RangeError");
    b->data.assign((size_t)l, 0);
    return Value::obj(b);
}
static Value buffer_slice(Interp& I, const Value& self, const Value* args, int argc)
{
    if (!self.is_obj() || self.o->kind != O_BUFFER) I.throw_error("this is not an ArrayBuffer", "TypeError");
    BufferObj* s = (BufferObj*)self.o;
    size_t b = rel_index(arg_int(I, ARG(0), 0), s->data.size());
    size_t e = rel_index(arg_int(I, ARG(1), (double)s->data.size()), s->data.size());
    if (e < b) e = b;
    BufferObj* r = new BufferObj();
    r->set_proto(I.buffer_proto);
    r->data.assign(s->data.begin() + b, s->data.begin() + e);
    return Value::obj(r);
}

// ---- Array
static ArrayObj* this_array(Interp& I, const Value& self)
{
    if (!self.is_obj() || self.o->kind != O_ARRAY) I.throw_error("this is not an array", "TypeError");
    return (ArrayObj*)self.o;
}
static Value array_ctor(Interp& I, const Value&, const Value* args, int argc)
{
    ArrayObj* a = I.new_array();
    Value av = Value::obj(a);
    if (argc == 1 && args[0].is_num()) {
        double l = args[0].n;
        if (l < 0 || l != std::floor(l) || l > 4294967295.0) I.throw_error("Invalid array length", "RangeError");
        a->el.resize((size_t)l);
    } else
        for (int i = 0; i < argc; ++i) a->el.push_back(args[i]);
    return av;
}
static Value array_push(Interp& I, const Value& self, const Value* args, int argc)
{
    ArrayObj* a = this_array(I, self);
    for (int i = 0; i < argc; ++i) a->el.push_back(args[i]);
    return Value((double)a->size());
}
static Value array_pop(Interp& I, const Value& self, const Value*, int)
{
    ArrayObj* a = this_array(I, self);
    if (a->size() == 0) return Value();
    Value v = a->el.back();
    a->el.pop_back();
    return v;
}
static Value array_shift(Interp& I, const Value& self, const Value*, int)
{
    ArrayObj* a = this_array(I, self);
    if (a->size() == 0) return Value();
    Value v = std::move(a->el[a->head]);
    a->el[a->head] = Value();
    a->head++;
    if (a->head > 64 && a->head * 2 > a->el.size()) {  // compact now and then
        a->el.erase(a->el.begin(), a->el.begin() + a->head);
        a->head = 0;
    }
    return v;
}
static Value array_unshift(Interp& I, const Value& self, const Value* args, int argc)
{
    ArrayObj* a = this_array(I, self);
    a->el.insert(a->el.begin() + a->head, args, args + argc);
    return Value((double)a->size());
}
static Value array_concat(Interp& I, const Value& self, const Value* args, int argc)
{
    ArrayObj* a = this_array(I, self);
    ArrayObj* r = I.new_array();
    Value rv = Value::obj(r);
    for (size_t i = 0; i < a->size(); ++i) r->el.push_back(a->at(i));
    for (int k = 0; k < argc; ++k) {
        if (args[k].is_obj() && args[k].o->kind == O_ARRAY) {
            ArrayObj* s = (ArrayObj*)args[k].o;
            for (size_t i = 0; i < s->size(); ++i) r->el.push_back(s->at(i));
        } else
            r->el.push_back(args[k]);
    }
    return rv;
}
static Value array_slice(Interp& I, const Value& self, const Value* args, int argc)
{
    ArrayObj* a = this_array(I, self);
    size_t b = rel_index(arg_int(I, ARG(0), 0), a->size());
    size_t e = rel_index(arg_int(I, ARG(1), (double)a->size()), a->size());
    ArrayObj* r = I.new_array();
    Value rv = Value::obj(r);
    for (size_t i = b; i < e; ++i) r->el.push_back(a->at(i));
    return rv;
}
static Value array_fill(Interp& I, const Value& self, const Value* args, int argc)
{
    ArrayObj* a = this_array(I, self);
    size_t b = rel_index(arg_int(I, ARG(1), 0), a->size());
    size_t e = rel_index(arg_int(I, ARG(2), (double)a->size()), a->size());
    for (size_t i = b; i < e; ++i) a->at(i) = ARG(0);
    return self;
}
static Value array_forEach(Interp& I, const Value& self, const Value* args, int argc)
{
    ArrayObj* a = this_array(I, self);
    for (size_t i = 0; i < a->size(); ++i) {
        Value cb_args[3] = {a->at(i), Value((double)i), self};
        I.call(ARG(0), Value(), cb_args, 3);
    }
    return Value();
}
static Value array_map(Interp& I, const Value& self, const Value* args, int argc)
{
    ArrayObj* a = this_array(I, self);
    ArrayObj* r = I.new_array();
    Value rv = Value::obj(r);
    for (size_t i = 0; i < a->size(); ++i) {
        Value cb_args[3] = {a->at(i), Value((double)i), self};
        r->el.push_back(I.call(ARG(0), Value(), cb_args, 3));
    }
    return rv;
}
static Value array_join(Interp& I, const Value& self, const Value* args, int argc)
{
    ArrayObj* a = this_array(I, self);
    U16 sep = ARG(0).is_undef() ? ascii(",") : I.to_string(ARG(0));
    U16 s;
    for (size_t i = 0; i < a->size(); ++i) {
        if (i) s += sep;
        const Value& e = a->at(i);
        if (e.tag != T_UNDEF && e.tag != T_NULL) s += I.to_string(e);
    }
    return Value::str(s);
}
static Value array_indexOf(Interp& I, const Value& self, const Value* args, int argc)
{
    ArrayObj* a = this_array(I, self);
    for (size_t i = 0; i < a->size(); ++i)
        if (I.strict_equals(a->at(i), ARG(0))) return Value((double)i);
    return Value(-1.0);
}
static Value array_isArray(Interp&, const Value&, const Value* args, int argc)
{
    return Value::boolean(argc > 0 && args[0].is_obj() && args[0].o->kind == O_ARRAY);
}

// ---- Object
static Value object_ctor(Interp& I, const Value&, const Value* args, int argc)
{
    if (argc > 0 && args[0].is_obj()) return args[0];
    return Value::obj(I.new_object());
}
static Value object_assign(Interp& I, const Value&, const Value* args, int argc)
{
    Value t = ARG(0);
    for (int k = 1; k < argc; ++k) {
        if (!args[k].is_obj()) continue;
        Obj* s = args[k].o;
        if (s->iprops)
            for (auto& kv : *s->iprops) I.set_index(t, Value((double)kv.first), kv.second);
        for (size_t i = 0; i < s->props.keys.size(); ++i) {
            Prop& p = s->props.vals[i];
            Value v = p.getter ? I.call_function((FuncObj*)p.getter, args[k], nullptr, 0) : p.v;
            I.set_prop(t, s->props.keys[i], v);
        }
    }
    return t;
}
static Value object_keys(Interp& I, const Value&, const Value* args, int argc)
{
    ArrayObj* r = I.new_array();
    Value rv = Value::obj(r);
    if (argc > 0 && args[0].is_obj()) {
        Obj* o = args[0].o;
        if (o->kind == O_ARRAY)
            for (size_t i = 0; i < ((ArrayObj*)o)->size(); ++i) r->el.push_back(Value::str(Interp::num_to_string((double)i)));
        if (o->iprops) {
            std::vector<uint32_t> ks;
            for (auto& kv : *o->iprops) ks.push_back(kv.first);
            std::sort(ks.begin(), ks.end());
            for (uint32_t k : ks) r->el.push_back(Value::str(Interp::num_to_string((double)k)));
        }
        for (Atom k : o->props.keys) r->el.push_back(Value::str(from_utf8(g_atoms.names[k])));
    }
    return rv;
}
static Value object_hasOwn(Interp& I, const Value& self, const Value* args, int argc)
{
    if (!self.is_obj()) return Value::boolean(false);
    uint32_t idx;
    Value k = ARG(0);
    if (I.key_to_index(k, idx)) {
        Obj* o = self.o;
        if (o->kind == O_ARRAY) return Value::boolean(idx < ((ArrayObj*)o)->size());
        if (o->kind == O_TYPED) return Value::boolean(idx < ((TypedObj*)o)->len);
        return Value::boolean(o->iprops && o->iprops->count(idx));
    }
    return Value::boolean(self.o->props.find(I.key_to_atom(k)) != nullptr);
}
static Value object_toString(Interp&, const Value&, const Value*, int) { return Value::str(ascii("[object Object]")); }

// ---- Math / Number / String
static Value math_floor(Interp& I, const Value&, const Value* args, int argc) { return Value(std::floor(I.to_number(ARG(0)))); }
static Value math_ceil(Interp& I, const Value&, const Value* args, int argc) { return Value(std::ceil(I.to_number(ARG(0)))); }
static Value math_round(Interp& I, const Value&, const Value* args, int argc) { return Value(std::floor(I.to_number(ARG(0)) + 0.5)); }
static Value math_trunc(Interp& I, const Value&, const Value* args, int argc) { return Value(std::trunc(I.to_number(ARG(0)))); }
static Value math_abs(Interp& I, const Value&, const Value* args, int argc) { return Value(std::fabs(I.to_number(ARG(0)))); }
static Value math_sqrt(Interp& I, const Value&, const Value* args, int argc) { return Value(std::sqrt(I.to_number(ARG(0)))); }
static Value math_log2(Interp& I, const Value&, const Value* args, int argc) { return Value(std::log2(I.to_number(ARG(0)))); }
static Value math_pow(Interp& I, const Value&, const Value* args, int argc) { return Value(std::pow(I.to_number(ARG(0)), I.to_number(ARG(1)))); }
static Value math_imul(Interp& I, const Value&, const Value* args, int argc)
{
    return Value((double)(int32_t)((uint32_t)I.to_int32(ARG(0)) * (uint32_t)I.to_int32(ARG(1))));
}
static Value math_min(Interp& I, const Value&, const Value* args, int argc)
{
    double r = INFINITY;
    for (int i = 0; i < argc; ++i) {
        double d = I.to_number(args[i]);
        if (std::isnan(d)) return Value(NAN);
        if (d < r) r = d;
    }
    return Value(r);
}
static Value math_max(Interp& I, const Value&, const Value* args, int argc)
{
    double r = -INFINITY;
    for (int i = 0; i < argc; ++i) {
        double d = I.to_number(args[i]);
        if (std::isnan(d)) return Value(NAN);
        if (d > r) r = d;
    }
    return Value(r);
}
static Value math_random(Interp&, const Value&, const Value*, int) { return Value((double)rand() / ((double)RAND_MAX + 1)); }
static Value number_fn(Interp& I, const Value&, const Value* args, int argc) { return Value(argc ? I.to_number(args[0]) : 0.0); }
static Value number_toString(Interp& I, const Value& self, const Value* args, int argc)
{
    double radix = ARG(0).is_undef() ? 10 : I.to_number(ARG(0));
    return Value::str(Interp::num_to_string(I.to_number(self), (int)radix));
}
static Value number_isInteger(Interp&, const Value&, const Value* args, int argc)
{
    return Value::boolean(argc > 0 && args[0].is_num() && std::isfinite(args[0].n) && args[0].n == std::floor(args[0].n));
}
static Value string_fn(Interp& I, const Value&, const Value* args, int argc) { return Value::str(argc ? I.to_string(args[0]) : U16()); }
static Value string_fromCharCode(Interp& I, const Value&, const Value* args, int argc)
{
    U16 s;
    for (int i = 0; i < argc; ++i) s += (char16_t)(Interp::to_uint32(I.to_number(args[i])) & 0xFFFF);
    return Value::str(s);
}
static Value string_charCodeAt(Interp& I, const Value& self, const Value* args, int argc)
{
    U16 s = I.to_string(self);
    double i = arg_int(I, ARG(0), 0);
    if (i < 0 || i >= (double)s.size()) return Value(NAN);
    return Value((double)s[(size_t)i]);
}
static Value string_charAt(Interp& I, const Value& self, const Value* args, int argc)
{
    U16 s = I.to_string(self);
    double i = arg_int(I, ARG(0), 0);
    if (i < 0 || i >= (double)s.size()) return Value::str(U16());
    return Value::str(U16(1, s[(size_t)i]));
}
static Value string_substring(Interp& I, const Value& self, const Value* args, int argc)
{
    U16 s = I.to_string(self);
    double a = arg_int(I, ARG(0), 0), b = arg_int(I, ARG(1), (double)s.size());
    if (a < 0) a = 0;
    if (b < 0) b = 0;
    if (a > (double)s.size()) a = (double)s.size();
    if (b > (double)s.size()) b = (double)s.size();
    if (a > b) std::swap(a, b);
    return Value::str(s.substr((size_t)a, (size_t)(b - a)));
}
static Value string_slice(Interp& I, const Value& self, const Value* args, int argc)
{
    U16 s = I.to_string(self);
    size_t b = rel_index(arg_int(I, ARG(0), 0), s.size());
    size_t e = rel_index(arg_int(I, ARG(1), (double)s.size()), s.size());
    if (e < b) e = b;
    return Value::str(s.substr(b, e - b));
}
static Value string_indexOf(Interp& I, const Value& self, const Value* args, int argc)
{
    U16 s = I.to_string(self), t = I.to_string(ARG(0));
    size_t p = s.find(t);
    return Value(p == U16::npos ? -1.0 : (double)p);
}
static Value string_toString(Interp& I, const Value& self, const Value*, int) { return Value::str(I.to_string(self)); }
static Value string_split(Interp& I, const Value& self, const Value* args, int argc)
{
    U16 s = I.to_string(self);
    ArrayObj* r = I.new_array();
    Value rv = Value::obj(r);
    if (ARG(0).is_undef()) {
        r->el.push_back(Value::str(s));
        return rv;
    }
    U16 sep = I.to_string(ARG(0));
    if (sep.empty()) {
        for (char16_t c : s) r->el.push_back(Value::str(U16(1, c)));
        return rv;
    }
    size_t at = 0;
    for (;;) {
        size_t f = s.find(sep, at);
        if (f == U16::npos) {
            r->el.push_back(Value::str(s.substr(at)));
            break;
        }
        r->el.push_back(Value::str(s.substr(at, f - at)));
        at = f + sep.size();
    }
    return rv;
}
static Value string_trim(Interp& I, const Value& self, const Value*, int)
{
    U16 s = I.to_string(self);
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == u' ' || s[a] == u'\t' || s[a] == u'\n' || s[a] == u'\r')) ++a;
    while (b > a && (s[b - 1] == u' ' || s[b - 1] == u'\t' || s[b - 1] == u'\n' || s[b - 1] == u'\r')) --b;
    return Value::str(s.substr(a, b - a));
}
static Value string_padStart(Interp& I, const Value& self, const Value* args, int argc)
{
    U16 s = I.to_string(self);
    size_t n = (size_t)arg_int(I, ARG(0), 0);
    U16 pad = ARG(1).is_undef() ? ascii(" ") : I.to_string(ARG(1));
    U16 r;
    while (!pad.empty() && r.size() + s.size() < n) r += pad[r.size() % pad.size()];
    return Value::str(r + s);
}

// ---- Error
static Value error_ctor_fn(Interp& I, const Value&, const Value* args, int argc)
{
    return I.make_error(ARG(0).is_undef() ? std::string() : narrow(I.to_string(ARG(0))), "Error");
}
static Value error_toString(Interp& I, const Value& self, const Value*, int)
{
    U16 s = I.to_string(I.get_prop(self, I.a_name));
    s += ascii(": ");
    s += I.to_string(I.get_prop(self, I.a_message));
    return Value::str(s);
}

// ---- Date (UTC)
static Value date_ctor(Interp& I, const Value&, const Value* args, int argc)
{
    DateObj* d = new DateObj();
    d->set_proto(I.date_proto);
    if (argc == 0) {
        d->ms = (double)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
    } else if (argc == 1) {
        d->ms = I.to_number(args[0]);
    } else {
        struct tm tmv;
        memset(&tmv, 0, sizeof tmv);
        tmv.tm_year = (int)I.to_number(ARG(0)) - 1900;
        tmv.tm_mon = (int)I.to_number(ARG(1));
        tmv.tm_mday = argc > 2 ? (int)I.to_number(ARG(2)) : 1;
        tmv.tm_hour = argc > 3 ? (int)I.to_number(ARG(3)) : 0;
        tmv.tm_min = argc > 4 ? (int)I.to_number(ARG(4)) : 0;
        tmv.tm_sec = argc > 5 ? (int)I.to_number(ARG(5)) : 0;
        d->ms = (double)timegm(&tmv) * 1000.0 + (argc > 6 ? I.to_number(ARG(6)) : 0);
    }
    return Value::obj(d);
}
static Value date_now(Interp&, const Value&, const Value*, int)
{
    return Value((double)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count());
}
static struct tm date_tm(Interp& I, const Value& self)
{
    if (!self.is_obj() || self.o->kind != O_DATE) I.throw_error("this is not a Date", "TypeError");
    time_t t = (time_t)std::floor(((DateObj*)self.o)->ms / 1000.0);
    struct tm r;
    gmtime_r(&t, &r);
    return r;
}
static Value date_getFullYear(Interp& I, const Value& s, const Value*, int) { return Value((double)date_tm(I, s).tm_year + 1900); }
static Value date_getMonth(Interp& I, const Value& s, const Value*, int) { return Value((double)date_tm(I, s).tm_mon); }
static Value date_getDate(Interp& I, const Value& s, const Value*, int) { return Value((double)date_tm(I, s).tm_mday); }
static Value date_getDay(Interp& I, const Value& s, const Value*, int) { return Value((double)date_tm(I, s).tm_wday); }
static Value date_getHours(Interp& I, const Value& s, const Value*, int) { return Value((double)date_tm(I, s).tm_hour); }
static Value date_getMinutes(Interp& I, const Value& s, const Value*, int) { return Value((double)date_tm(I, s).tm_min); }
static Value date_getSeconds(Interp& I, const Value& s, const Value*, int) { return Value((double)date_tm(I, s).tm_sec); }
static Value date_getTime(Interp& I, const Value& s, const Value*, int)
{
    if (!s.is_obj() || s.o->kind != O_DATE) I.throw_error("this is not a Date", "TypeError");
    return Value(((DateObj*)s.o)->ms);
}

// ---- TextEncoder / TextDecoder
static Value text_coder_ctor(Interp& I, const Value&, const Value*, int) { return Value::obj(I.new_object()); }
static Value text_encode(Interp& I, const Value&, const Value* args, int argc)
{
    std::string s = narrow(I.to_string(ARG(0)));
    TypedObj* t = I.new_typed(E_U8, s.size());
    memcpy(t->ptr(), s.data(), s.size());
    return Value::obj(t);
}
static Value text_decode(Interp& I, const Value&, const Value* args, int argc)
{
    Value a = ARG(0);
    if (!a.is_obj() || a.o->kind != O_TYPED) return Value::str(U16());
    TypedObj* t = (TypedObj*)a.o;
    return Value::str(from_utf8(std::string((const char*)t->ptr(), t->len * kElemSize[t->ek])));
}

// ---- host
static Value host_print(Interp& I, const Value&, const Value* args, int argc)
{
    for (int i = 0; i < argc; ++i) {
        if (i) fputc(' ', stdout);
        fputs(narrow(I.to_string(args[i])).c_str(), stdout);
    }
    fputc('\n', stdout);
    fflush(stdout);
    return Value();
}
static Value host_readFile(Interp& I, const Value&, const Value* args, int argc)
{
    std::string path = narrow(I.to_string(ARG(0)));
    std::ifstream f(path, std::ios::binary);
    if (!f) I.throw_error("readFile: cannot open " + path);
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    TypedObj* t = I.new_typed(E_U8, data.size());
    memcpy(t->ptr(), data.data(), data.size());
    return Value::obj(t);
}
static Value host_writeFile(Interp& I, const Value&, const Value* args, int argc)
{
    std::string path = narrow(I.to_string(ARG(0)));
    std::ofstream f(path, std::ios::binary);
    if (!f) I.throw_error("writeFile: cannot open " + path);
    Value d = ARG(1);
    if (d.is_obj() && d.o->kind == O_TYPED) {
        TypedObj* t = (TypedObj*)d.o;
        f.write((const char*)t->ptr(), (std::streamsize)(t->len * kElemSize[t->ek]));
    } else {
        std::string s = narrow(I.to_string(d));
        f.write(s.data(), (std::streamsize)s.size());
    }
    return Value();
}
static Value host_clock(Interp&, const Value&, const Value*, int)
{
    return Value(std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count());
}

void Interp::setup()
{
    a_length = A("length");
    a_prototype = A("prototype");
    a_constructor = A("constructor");
    a_message = A("message");
    a_name = A("name");
    a_buffer = A("buffer");
    a_byteOffset = A("byteOffset");
    a_byteLength = A("byteLength");
    a_BYTES_PER_ELEMENT = A("BYTES_PER_ELEMENT");
    a_stack = A("stack");

    object_proto = new Obj(O_PLAIN);
    obj_retain(object_proto);
    auto mkproto = [&]() {
        Obj* p = new Obj(O_PLAIN);
        obj_retain(p);
        p->set_proto(object_proto);
        return p;
    };
    function_proto = mkproto();
    array_proto = mkproto();
    error_proto = mkproto();
    date_proto = mkproto();
    buffer_proto = mkproto();
    string_proto = mkproto();
    number_proto = mkproto();
    global = new_object();
    obj_retain(global);
    Value gv = Value::obj(global);
    for (const char* nm : {"globalThis", "window", "self", "global"}) global->props.insert(A(nm))->v = gv;
    global->props.insert(A("NaN"))->v = Value(NAN);
    global->props.insert(A("Infinity"))->v = Value(INFINITY);

    auto def_ctor = [&](const char* name, NativeFn fn, Obj* proto, int kind = 0) {
        FuncObj* f = make_native(name, fn, kind);
        Value fv = Value::obj(f);
        f->props.insert(a_prototype)->v = Value::obj(proto);
        proto->props.insert(a_constructor)->v = fv;
        global->props.insert(A(name))->v = fv;
        return f;
    };

    // Object
    FuncObj* objc = def_ctor("Object", object_ctor, object_proto);
    def_native(objc, "assign", object_assign);
    def_native(objc, "keys", object_keys);
    def_native(object_proto, "hasOwnProperty", object_hasOwn);
    def_native(object_proto, "toString", object_toString);
    def_ctor("Function", object_ctor, function_proto);

    // Array
    FuncObj* arrc = def_ctor("Array", array_ctor, array_proto);
    def_native(arrc, "isArray", array_isArray);
    def_native(array_proto, "push", array_push);
    def_native(array_proto, "pop", array_pop);
    def_native(array_proto, "shift", array_shift);
    def_native(array_proto, "unshift", array_unshift);
    def_native(array_proto, "concat", array_concat);
    def_native(array_proto, "slice", array_slice);
    def_native(array_proto, "fill", array_fill);
    def_native(array_proto, "forEach", array_forEach);
    def_native(array_proto, "map", array_map);
    def_native(array_proto, "join", array_join);
    def_native(array_proto, "indexOf", array_indexOf);

    // typed arrays
    static const char* const tnames[6] = {"Uint8Array", "Uint16Array", "Uint32Array", "Int8Array", "Int16Array", "Int32Array"};
    for (int k = 0; k < 6; ++k) {
        typed_proto[k] = mkproto();
        typed_ctor[k] = def_ctor(tnames[k], mjs::typed_ctor, typed_proto[k], k);
        typed_ctor[k]->props.insert(a_BYTES_PER_ELEMENT)->v = Value((double)kElemSize[k]);
        def_native(typed_proto[k], "set", typed_set_m);
        def_native(typed_proto[k], "subarray", typed_subarray);
        def_native(typed_proto[k], "slice", typed_slice);
        def_native(typed_proto[k], "fill", typed_fill);
        def_native(typed_proto[k], "map", typed_map);
        def_native(typed_proto[k], "forEach", typed_forEach);
        def_native(typed_proto[k], "join", typed_join);
        def_native(typed_proto[k], "indexOf", typed_indexOf);
    }
    def_ctor("ArrayBuffer", buffer_ctor, buffer_proto);
    def_native(buffer_proto, "slice", buffer_slice);

    // Math
    Obj* math = new_object();
    global->props.insert(A("Math"))->v = Value::obj(math);
    def_native(math, "floor", math_floor);
    def_native(math, "ceil", math_ceil);
    def_native(math, "round", math_round);
    def_native(math, "trunc", math_trunc);
    def_native(math, "abs", math_abs);
    def_native(math, "sqrt", math_sqrt);
    def_native(math, "log2", math_log2);
    def_native(math, "pow", math_pow);
    def_native(math, "imul", math_imul);
    def_native(math, "min", math_min);
    def_native(math, "max", math_max);
    def_native(math, "random", math_random);
    math->props.insert(A("PI"))->v = Value(3.141592653589793);

    // Number / String
    FuncObj* numc = def_ctor("Number", number_fn, number_proto);
    numc->props.insert(A("POSITIVE_INFINITY"))->v = Value(INFINITY);
    numc->props.insert(A("NEGATIVE_INFINITY"))->v = Value(-INFINITY);
    numc->props.insert(A("MAX_SAFE_INTEGER"))->v = Value(9007199254740991.0);
    numc->props.insert(A("MAX_VALUE"))->v = Value(1.7976931348623157e308);
    numc->props.insert(A("NaN"))->v = Value(NAN);
    def_native(numc, "isInteger", number_isInteger);
    def_native(number_proto, "toString", number_toString);
    FuncObj* strc = def_ctor("String", string_fn, string_proto);
    def_native(strc, "fromCharCode", string_fromCharCode);
    def_native(string_proto, "charCodeAt", string_charCodeAt);
    def_native(string_proto, "charAt", string_charAt);
    def_native(string_proto, "substring", string_substring);
    def_native(string_proto, "slice", string_slice);
    def_native(string_proto, "indexOf", string_indexOf);
    def_native(string_proto, "toString", string_toString);
    def_native(string_proto, "padStart", string_padStart);
    def_native(string_proto, "split", string_split);
    def_native(string_proto, "trim", string_trim);

    // Error (+ the subclasses the interpreter itself throws)
    error_ctor = def_ctor("Error", error_ctor_fn, error_proto);
    error_proto->props.insert(a_name)->v = Value::str(ascii("Error"));
    error_proto->props.insert(a_message)->v = Value::str(U16());
    def_native(error_proto, "toString", error_toString);
    for (const char* nm : {"TypeError", "RangeError", "ReferenceError", "SyntaxError"}) global->props.insert(A(nm))->v = Value::obj(error_ctor);

    // Date
    FuncObj* datec = def_ctor("Date", date_ctor, date_proto);
    def_native(datec, "now", date_now);
    def_native(date_proto, "getFullYear", date_getFullYear);
    def_native(date_proto, "getMonth", date_getMonth);
    def_native(date_proto, "getDate", date_getDate);
    def_native(date_proto, "getDay", date_getDay);
    def_native(date_proto, "getHours", date_getHours);
    def_native(date_proto, "getMinutes", date_getMinutes);
    def_native(date_proto, "getSeconds", date_getSeconds);
    def_native(date_proto, "getTime", date_getTime);
    def_native(date_proto, "valueOf", date_getTime);

    // TextEncoder / TextDecoder
    Obj* tep = mkproto();
    def_ctor("TextEncoder", text_coder_ctor, tep);
    def_native(tep, "encode", text_encode);
    Obj* tdp = mkproto();
    def_ctor("TextDecoder", text_coder_ctor, tdp);
    def_native(tdp, "decode", text_decode);

    // console + host
    Obj* console = new_object();
    global->props.insert(A("console"))->v = Value::obj(console);
    def_native(console, "log", host_print);
    def_native(console, "error", host_print);
    def_native(global, "print", host_print);
    def_native(global, "readFile", host_readFile);
    def_native(global, "writeFile", host_writeFile);
    def_native(global, "clock", host_clock);
    def_native(global, "Boolean", [](Interp& I, const Value&, const Value* args, int argc) {
        return Value::boolean(argc > 0 && I.to_bool(args[0]));
    });
    def_native(global, "isNaN", [](Interp& I, const Value&, const Value* args, int argc) {
        return Value::boolean(std::isnan(I.to_number(ARG(0))));
    });
    def_native(global, "parseInt", [](Interp& I, const Value&, const Value* args, int argc) {
        std::string s = narrow(I.to_string(ARG(0)));
        int radix = ARG(1).is_undef() ? 10 : (int)I.to_number(ARG(1));
        char* end;
        long long v = strtoll(s.c_str(), &end, radix);
        if (end == s.c_str()) return Value(NAN);
        return Value((double)v);
    });
}

// text_coder objects need their prototype: construct() of a native passes the constructor as `this`
static Value text_coder_ctor_fix(Interp& I, const Value& self, const Value*, int)
{
    Obj* o = I.new_object();
    if (self.is_obj() && self.o->kind == O_FUNC) {
        Value pv = I.get_prop(self, I.a_prototype);
        if (pv.is_obj()) o->set_proto(pv.o);
    }
    return Value::obj(o);
}

// =====================================================================================================
// API
// =====================================================================================================
Interp* create()
{
    Interp* I = new Interp();
    I->setup();
    // TextEncoder / TextDecoder instances must inherit encode / decode
    for (const char* nm : {"TextEncoder", "TextDecoder"}) {
        Prop* p = I->global->props.find(A(nm));
        ((FuncObj*)p->v.o)->native = text_coder_ctor_fix;
    }
    return I;
}
void destroy(Interp* I) { delete I; }

void set_args(Interp* I, const std::vector<std::string>& args)
{
    ArrayObj* a = I->new_array();
    for (auto& s : args) a->el.push_back(Value::str(from_utf8(s)));
    I->global->props.insert(A("scriptArgs"))->v = Value::obj(a);
}

bool run_source(Interp* I, const std::string& src, const std::string& name, std::string& err)
{
    try {
        Parser p;
        p.lx.src = src;
        p.lx.file = name;
        // "use strict" directive and friends are plain expression statements
        FuncNode* prog = p.parse_program();
        resolve_func(prog);
        FuncObj* f = new FuncObj();
        Value fv = Value::obj(f);
        f->set_proto(I->function_proto);
        f->node = prog;
        I->call_function(f, Value(), nullptr, 0);
        return true;
    } catch (JsThrow& t) {
        try {
            err = "uncaught " + narrow(I->to_string(t.v));
        } catch (...) {
            err = "uncaught exception";
        }
        return false;
    } catch (std::exception& e) {
        err = e.what();
        return false;
    }
}

bool run_file(Interp* I, const std::string& path, std::string& err)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        err = "cannot open " + path;
        return false;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    return run_source(I, ss.str(), path, err);
}

}  // namespace mjs

#ifndef MINIJS_NO_MAIN
// usage: minijs script.js [script2.js ...] [-- args...]
int main(int argc, char** argv)
{
    std::vector<std::string> scripts, args;
    bool rest = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (!rest && a == "--") {
            rest = true;
            continue;
        }
        (rest ? args : scripts).push_back(a);
    }
    if (scripts.empty()) {
        fprintf(stderr, "usage: minijs script.js [more.js ...] [-- args...]\n");
        return 2;
    }
    mjs::Interp* I = mjs::create();
    mjs::set_args(I, args);
    for (auto& s : scripts) {
        std::string err;
        if (!mjs::run_file(I, s, err)) {
            fprintf(stderr, "%s: %s\n", s.c_str(), err.c_str());
            return 1;
        }
    }
    return 0;
}
#endif
