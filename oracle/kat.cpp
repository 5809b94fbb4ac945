// ORACLE known-answer tests: the reference's own unit tests, transcribed literal-for-literal.
// Every `assert_eq!` on f32 in the reference is a bit-exact EQ here; every
// `assert_relative_eq!(.., epsilon = COLLISION_EPSILON)` uses the same approx-0.3 predicate.
// Source of each block is cited (file:line under /root/reference).  These vectors are what
// pins the cgmath-0.17 restatement in cgm.hpp.
#include <cstdio>
#include <vector>
#include "simplex.hpp"
#include "world.hpp"
#include "compound.hpp"

using namespace mgfo;

static int g_fail = 0, g_pass = 0;
#define CHECK(cond) do { if (cond) ++g_pass; else { ++g_fail; std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)
#define EQ(a, b) do { float _a = (a), _b = (b); if (_a == _b) ++g_pass; else { ++g_fail; std::printf("FAIL %s:%d: %s == %s (%.9g vs %.9g)\n", __FILE__, __LINE__, #a, #b, _a, _b); } } while (0)
#define EQV(v, X, Y, Z) do { Vec3 _v = (v); if (_v.x == (X) && _v.y == (Y) && _v.z == (Z)) ++g_pass; else { ++g_fail; std::printf("FAIL %s:%d: %s = (%.9g %.9g %.9g)\n", __FILE__, __LINE__, #v, _v.x, _v.y, _v.z); } } while (0)
#define REL(a, b, eps) do { float _a = (a), _b = (b); if (relative_eq(_a, _b, eps)) ++g_pass; else { ++g_fail; std::printf("FAIL %s:%d: %s ~ %s (%.9g vs %.9g)\n", __FILE__, __LINE__, #a, #b, _a, _b); } } while (0)
#define RELV(v, X, Y, Z, eps) do { Vec3 _v = (v); if (relative_eq(_v.x, (X), eps) && relative_eq(_v.y, (Y), eps) && relative_eq(_v.z, (Z), eps)) ++g_pass; else { ++g_fail; std::printf("FAIL %s:%d: %s = (%.9g %.9g %.9g)\n", __FILE__, __LINE__, #v, _v.x, _v.y, _v.z); } } while (0)

static const float E = COLLISION_EPSILON;

static void test_ray_intersections() {  // collision.rs:1543-1637
    Intersection it;
    Capsule c{v3(0, 0, 0), v3(1, 0, 0), 1.0f};
    Ray r{v3(1, -3, 0), normalize(v3(-0.25f, 1, 0))};
    CHECK(intersection(r, c, &it));
    RELV(it.p, 0.5f, -1.0f, 0.0f, E);
    RELV(r.p + r.d * it.t, 0.5f, -1.0f, 0.0f, E);
    r = Ray{v3(0, -3, 0), normalize(v3(0.25f, 1, 0))};
    CHECK(intersection(r, c, &it));
    RELV(it.p, 0.5f, -1.0f, 0.0f, E);
    RELV(r.p + r.d * it.t, 0.5f, -1.0f, 0.0f, E);
    c = Capsule{v3(0, 0, 0), v3(0, 2, 0), 2.0f};
    r = Ray{v3(4, 1, 0), v3(-1, 0, 0)};
    CHECK(intersection(r, c, &it));
    EQV(it.p, 2.0f, 1.0f, 0.0f); EQ(it.t, 2.0f);
    c = Capsule{v3(0, 0, 0), v3(1, 0, 0), 1.0f};
    r = Ray{v3(3, 0, 0), v3(-1, 0, 0)};
    CHECK(intersection(r, c, &it));
    EQV(it.p, 2.0f, 0.0f, 0.0f); EQ(it.t, 1.0f);
    r = Ray{v3(-2, 0, 0), v3(1, 0, 0)};
    CHECK(intersection(r, c, &it));
    EQV(it.p, -1.0f, 0.0f, 0.0f); EQ(it.t, 1.0f);
    r = Ray{v3(-2, 0.5f, 0), v3(1, 0, 0)};
    CHECK(intersection(r, c, &it));
    RELV(it.p, -0.8660254037844386f, 0.5f, 0.0f, F32_EPSILON);
    REL(it.t, 1.13397459621556196f, E);
    r = Ray{v3(3, 0.5f, 0), v3(-1, 0, 0)};
    CHECK(intersection(r, c, &it));
    RELV(it.p, 1.8660254037844386f, 0.5f, 0.0f, F32_EPSILON);
    REL(it.t, 1.13397459621556196f, E);
}

static void test_sphere_penetration() {  // collision.rs:1647-1672
    Sphere s1{v3(0, 0, 0), 1.0f}, s2{v3(2, 0, 0), 1.5f};
    float sep;
    CHECK(!separation(s1, s2, &sep));
    CHECK(!separation(s2, s1, &sep));
    s2 = Sphere{v3(2, 0, 0), 0.75f};
    CHECK(separation(s1, s2, &sep));
    EQ(sep, 0.25f);
}

static void test_moving_spheres_collision() {  // collision.rs:1675-1696
    Moving<Sphere> s1{Sphere{v3(-3, 0, 0), 1.0f}, v3(1, 0, 0)};
    Moving<Sphere> s2{Sphere{v3(3, 0, 0), 2.0f}, v3(-2, 0, 0)};
    Contact col{}; bool any = false;
    moving_contacts_moving(s1, s2, [&](const Contact& c) { col = c; any = true; });
    CHECK(any);
    EQ(col.t, 1.0f);
    EQV(col.a, -1.0f, 0.0f, 0.0f); EQV(col.b, -1.0f, 0.0f, 0.0f); EQV(col.n, 1.0f, 0.0f, 0.0f);
}

static void test_sphere_rect_collision() {  // collision.rs:1699-1758
    Rectangle floor{v3(0, 1, 0), {v3(1, 0, 0), v3(0, 0, 1)}, {3.0f, 3.0f}};
    Moving<Sphere> sc{Sphere{v3(0, 13, 0), 2.0f}, v3(0, -10, 0)};
    int calls = 0;
    CHECK(contacts(floor, sc, [&](const Contact& c) {
        ++calls; EQV(c.a, 0.0f, 1.0f, 0.0f); EQV(c.b, 0.0f, 1.0f, 0.0f); EQ(c.t, 1.0f); EQV(c.n, 0.0f, 1.0f, 0.0f);
    }));
    CHECK(moving_contacts_poly(sc, floor, [&](const Contact& c) {
        ++calls; EQV(c.a, 0.0f, 1.0f, 0.0f); EQV(c.b, 0.0f, 1.0f, 0.0f); EQ(c.t, 1.0f); EQV(c.n, 0.0f, -1.0f, 0.0f);
    }));
    Moving<Sphere> sc2{Sphere{v3(0, 13, 0), 2.0f}, v3(0, -20, 0)};
    CHECK(contacts(floor, sc2, [&](const Contact& c) {
        ++calls; EQV(c.a, 0.0f, 1.0f, 0.0f); EQV(c.b, 0.0f, 1.0f, 0.0f); EQ(c.t, 0.5f); EQV(c.n, 0.0f, 1.0f, 0.0f);
    }));
    Moving<Sphere> scc{Sphere{v3(0, 13, 0), 2.0f}, v3(0, -10, 3)};
    CHECK(contacts(floor, scc, [&](const Contact& c) {
        ++calls; EQV(c.a, 0.0f, 1.0f, 3.0f); EQV(c.b, 0.0f, 1.0f, 3.0f); EQ(c.t, 1.0f); EQV(c.n, 0.0f, 1.0f, 0.0f);
    }));
    CHECK(calls == 4);
    Moving<Sphere> smc{Sphere{v3(0, 13, 0), 2.0f}, v3(0, -10, 3.00001f)};
    CHECK(!contacts(floor, smc, [&](const Contact&) {}));
}

static void test_sphere_tri_collision() {  // collision.rs:1761-1814
    Triangle floor{v3(1, 1, 0), v3(0, 1, -1), v3(0, 1, 1)};  // a, b, c
    int calls = 0;
    Moving<Sphere> sc{Sphere{v3(0, 13, 0), 2.0f}, v3(0, -10, 0)};
    CHECK(contacts(floor, sc, [&](const Contact& c) {
        ++calls; EQV(c.a, 0.0f, 1.0f, 0.0f); EQV(c.b, 0.0f, 1.0f, 0.0f); EQ(c.t, 1.0f); EQV(c.n, 0.0f, 1.0f, 0.0f);
    }));
    Moving<Sphere> scc{Sphere{v3(0, 13, 0), 2.0f}, v3(0, -10, 1)};
    CHECK(contacts(floor, scc, [&](const Contact& c) {
        ++calls; RELV(c.a, 0.0f, 1.0f, 1.0f, E); RELV(c.b, 0.0f, 1.0f, 1.0f, E);
        CHECK((1.0f - c.t) < E); EQV(c.n, 0.0f, 1.0f, 0.0f);
    }));
    Moving<Sphere> smc{Sphere{v3(0, 13, 0), 2.0f}, v3(0, -10, 1.00001f)};
    CHECK(!contacts(floor, smc, [&](const Contact&) { CHECK(false); }));
    Moving<Sphere> sce{Sphere{v3(0, 13, 0), 2.0f}, v3(0.5f, -10, 0.5f)};
    CHECK(contacts(floor, sce, [&](const Contact& c) {
        ++calls; EQV(c.a, 0.5f, 1.0f, 0.5f); EQV(c.b, 0.5f, 1.0f, 0.5f); EQ(c.t, 1.0f); EQV(c.n, 0.0f, 1.0f, 0.0f);
    }));
    CHECK(calls == 3);
}

static void test_obb_collision() {  // collision.rs:1823-1843
    OBB box1{v3(0, 0, 0), quat_one(), v3(1, 1, 1)};
    OBB box2{v3(0, 1, 0), quat_one(), v3(1, 1.5f, 1)};
    Contact col;
    CHECK(gjk_contact(box1, box2, &col));
    EQ(col.a.y, 1.0f); EQ(col.b.y, -0.5f);
    CHECK(gjk_contact(box2, box1, &col));
    EQ(col.b.y, 1.0f); EQ(col.a.y, -0.5f);
    OBB box3{v3(0, 4.1f, 0), quat_one(), v3(1, 1.5f, 1)};
    CHECK(!gjk_contact(box1, box3, &col));
    OBB box4{v3(0, 2, 0), from_arc(v3(1, 0, 0), v3(0, 1, 0)), v3(1.7f, 1.5f, 1)};
    CHECK(gjk_contact(box1, box4, &col));
    EQ(col.a.y, 1.0f); EQ(col.b.y, 0.30000007f);
}

static void test_capsule_moving_sphere() {  // collision.rs:1853-1874
    Capsule c{v3(4, 3, 5.5f), v3(0, 1, 0), 2.0f};
    Moving<Sphere> s{Sphere{v3(0, 3, 5.5f), 1.0f}, v3(1, 0, 0)};
    Contact col;
    CHECK(last_contact(c, s, &col));
    EQ(col.t, 1.0f); EQV(col.a, 2.0f, 3.0f, 5.5f); EQV(col.b, 2.0f, 3.0f, 5.5f);
    bool any = false;
    moving_contacts_static(s, c, [&](const Contact& k) { col = k; any = true; });  // collision.rs:1870 -> :1368
    CHECK(any);
    EQ(col.t, 1.0f); EQV(col.a, 2.0f, 3.0f, 5.5f); EQV(col.b, 2.0f, 3.0f, 5.5f);
}

static void test_moving_capsule_collision() {  // collision.rs:1877-1980
    Contact col;
    Capsule s{v3(4, 3, 5.5f), v3(0, 1, 0), 2.0f};
    Moving<Capsule> c{Capsule{v3(0, 3, 5.5f), v3(0, 1, 0), 1.0f}, v3(1, 0, 0)};
    CHECK(last_contact(s, c, &col));
    EQ(col.t, 1.0f); EQV(col.a, 2.0f, 3.5f, 5.5f); EQV(col.b, 2.0f, 3.5f, 5.5f);
    s = Capsule{v3(4, 3, 5.5f), v3(0, 1, 0), 1.0f};
    c = Moving<Capsule>{Capsule{v3(0, 3, 5.5f), v3(0, 1, 0), 2.0f}, v3(1, 0, 0)};
    CHECK(last_contact(s, c, &col));
    EQV(col.a, 3.0f, 3.5f, 5.5f); EQV(col.b, 3.0f, 3.5f, 5.5f); EQ(col.t, 1.0f);
    s = Capsule{v3(1, 0, 0), v3(1, 0, 0), 1.0f};
    c = Moving<Capsule>{Capsule{v3(-2, 0, 0), v3(-1, 0, 0), 1.0f}, v3(2, 0, 0)};
    CHECK(last_contact(s, c, &col));
    EQV(col.a, 0.0f, 0.0f, 0.0f); EQV(col.b, 0.0f, 0.0f, 0.0f); EQ(col.t, 0.5f);
    s = Capsule{v3(0, 0, 0), v3(1, 0, 0), 1.0f};
    c = Moving<Capsule>{Capsule{v3(0, 0, 0), v3(-1, 0, 0), 1.0f}, v3(2, 0, 0)};
    CHECK(last_contact(s, c, &col));
    EQV(col.a, -1.0f, 0.0f, 0.0f); EQV(col.b, 1.0f, 0.0f, 0.0f); EQ(col.t, 0.0f);
    s = Capsule{v3(4, 3, 5.5f), v3(0, 1, 0), 2.0f};
    c = Moving<Capsule>{Capsule{v3(0, 2, 5.5f), v3(0, 1, 0), 1.0f}, v3(1, 0, 0)};
    CHECK(last_contact(s, c, &col));
    EQ(col.t, 1.0f); EQV(col.a, 2.0f, 3.0f, 5.5f); EQV(col.b, 2.0f, 3.0f, 5.5f);
    c = Moving<Capsule>{Capsule{v3(0, 2.5f, 5.5f), v3(0, 1, 0), 1.0f}, v3(1, 0, 0)};
    CHECK(last_contact(s, c, &col));
    EQ(col.t, 1.0f); EQV(col.a, 2.0f, 3.25f, 5.5f); EQV(col.b, 2.0f, 3.25f, 5.5f);
}

static void test_capsule_rect_collision() {  // collision.rs:1983-2003
    Rectangle floor{v3(0, 1, 0), {v3(1, 0, 0), v3(0, 0, 1)}, {3.0f, 3.0f}};
    Moving<Capsule> cap{Capsule{v3(1, 13, 0), v3(3, 0, 0), 2.0f}, v3(0, -10, 0)};
    std::vector<Contact> cs;
    contacts(floor, cap, [&](const Contact& c) { cs.push_back(c); });
    CHECK(cs.size() >= 2);
    if (cs.size() >= 2) {
        EQ(cs[0].t, 1.0f);
        RELV(cs[0].a, 1.0f, 1.0f, 0.0f, E);
        RELV(cs[1].a, 3.0f, 1.0f, 0.0f, E);
    }
}

static void test_capsule_tri_collision() {  // collision.rs:2006-2268
    Triangle floor{v3(1, 1, 0), v3(0, 1, -1), v3(0, 1, 1)};
    std::vector<Contact> cs;
    auto run = [&](Vec3 a, Vec3 d, float r, Vec3 v) {
        contacts(floor, Moving<Capsule>{Capsule{a, d, r}, v}, [&](const Contact& c) { cs.push_back(c); });
    };
    auto last = [&](Vec3 a, Vec3 d, float r, Vec3 v, Contact* out) {
        return last_contact(floor, Moving<Capsule>{Capsule{a, d, r}, v}, out);
    };
    // capsule_clip_edge (:2012)
    run(v3(0.9f, 3, 1), v3(0, 0, -2), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() >= 2);
    if (cs.size() >= 2) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.9f, 1.0f, 0.1f, E); RELV(cs[1].a, 0.9f, 1.0f, -0.1f, E); }
    cs.clear();
    // capsule_clip_off_center (:2026)
    run(v3(0.9f, 3, 0), v3(0, 0, 2), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() >= 2);
    if (cs.size() >= 2) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.9f, 1.0f, 0.0f, E); RELV(cs[1].a, 0.9f, 1.0f, 0.1f, E); }
    cs.clear();
    // (:2039)
    run(v3(0.9f, 3, 0), v3(0, 0, -2), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() >= 2);
    if (cs.size() >= 2) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.9f, 1.0f, 0.0f, E); RELV(cs[1].a, 0.9f, 1.0f, -0.1f, E); }
    cs.clear();
    // capsule_through_center (:2052)
    run(v3(0.9f, 2, 0), v3(1, 0, 0), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() >= 2);
    if (cs.size() >= 2) { EQ(cs[0].t, 0.0f); RELV(cs[0].a, 0.9f, 1.0f, 0.0f, E); RELV(cs[1].a, 1.0f, 1.0f, 0.0f, E); }
    cs.clear();
    // capsule_tilted_center (:2065)
    Contact col;
    CHECK(last(v3(0.5f, 4, 0), v3(-1, -0.5f, 0), 1.0f, v3(0, -2, 0), &col));
    EQ(col.t, 0.81598306f);
    RELV(col.a, 0.0f, 1.0f, 0.0f, E);
    // (:2093)
    CHECK(last(v3(0.5f, 4, 0), v3(-1, -1, 2), 1.0f, v3(0, -2, 0), &col));
    RELV(col.a, 0.0f, 1.0f, 1.0f, E);
    EQ(col.t, 0.7022774f);
    // capsule_parallel_to_edge (:2104)
    run(v3(-1, 2, 2), v3(0, 0, -2), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 2);
    if (cs.size() == 2) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.0f, 1.0f, 1.0f, E); RELV(cs[1].a, 0.0f, 1.0f, 0.0f, E); }
    cs.clear();
    // (:2118)
    run(v3(-1, 4, 2), v3(0, -2, -2), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 1);
    if (cs.size() >= 1) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.0f, 1.0f, 0.0f, E); }
    // (:2130) NOTE: the reference does not clear `contacts` here, then asserts len()==1:
    // i.e. this call must emit NO contact.
    run(v3(-1, 4, 0), v3(0, 2, -2), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 1);
    if (cs.size() >= 1) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.0f, 1.0f, 0.0f, E); }
    cs.clear();
    // (:2143)
    run(v3(-1, 2, 2), v3(0, 0, -4), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 2);
    if (cs.size() == 2) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.0f, 1.0f, 1.0f, E); RELV(cs[1].a, 0.0f, 1.0f, -1.0f, E); }
    cs.clear();
    // (:2157)
    run(v3(-1, 2, -2), v3(0, 0, 4), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 2);
    if (cs.size() == 2) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.0f, 1.0f, -1.0f, E); RELV(cs[1].a, 0.0f, 1.0f, 1.0f, E); }
    cs.clear();
    // new floor (:2171)
    floor = Triangle{v3(1, 1, 0), v3(0, 1, 2), v3(0, 1, -2)};
    run(v3(-0.5f, 2, 0.5f), v3(0, 0, -1), 0.5f, v3(0, -1, 0));
    CHECK(cs.size() == 2);
    if (cs.size() == 2) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.0f, 1.0f, 0.5f, E); RELV(cs[1].a, 0.0f, 1.0f, -0.5f, E); }
    cs.clear();
    // capsule_perp_to_edge (:2190)
    run(v3(-1, 2, 0), v3(-3, 0, 0), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 1);
    if (cs.size() >= 1) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.0f, 1.0f, 0.0f, E); }
    cs.clear();
    // (:2203)
    run(v3(-4, 2, 0), v3(3, 0, 0), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 1);
    if (cs.size() >= 1) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 0.0f, 1.0f, 0.0f, E); }
    cs.clear();
    // capsule_next_to_vert (:2216)
    run(v3(2, 2, 1), v3(0, 0, -2), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 1);
    if (cs.size() >= 1) { EQ(cs[0].t, 1.0f); RELV(cs[0].a, 1.0f, 1.0f, 0.0f, E); }
    cs.clear();
    // capsule_next_to_vert_skewed (:2229)
    run(v3(2, 2, 1), v3(0, -1, -2), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 1);
    if (cs.size() >= 1) { EQ(cs[0].t, 0.5f); RELV(cs[0].a, 1.0f, 1.0f, 0.0f, E); }
    cs.clear();
    // capsule_intersects_tri_plane (:2242)
    run(v3(0, 4, 0), v3(-2, -4, 0), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 1);
    if (cs.size() >= 1) { REL(cs[0].t, 0.7639319f, F32_EPSILON); RELV(cs[0].a, 0.0f, 1.0f, 0.0f, E); }
    cs.clear();
    // (:2255)
    run(v3(-1, 2, 0), v3(-1, -2, 0), 1.0f, v3(0, -1, 0));
    CHECK(cs.size() == 1);
    if (cs.size() >= 1) { REL(cs[0].t, 1.0f, F32_EPSILON); RELV(cs[0].a, 0.0f, 1.0f, 0.0f, E); }
    cs.clear();
}

static void test_bvh() {  // bvh.rs:514-529
    Sphere a{v3(0, 5, 0), 1.0f}, b{v3(0, 8, 0), 1.0f}, c{v3(3, 0, 0), 1.0f};
    BVH<size_t> bvh;
    bvh.insert(bounds(a), 1); bvh.insert(bounds(b), 2); bvh.insert(bounds(c), 3);
    size_t found = 0;
    bvh.query(bounds(a), [&](const size_t& id) { ++found; CHECK(id == 1); });
    bvh.query(bounds(b), [&](const size_t& id) { ++found; CHECK(id == 2); });
    bvh.query(bounds(c), [&](const size_t& id) { ++found; CHECK(id == 3); });
    CHECK(found == 3);
}

static void test_aabb() {  // bounds.rs:330-352
    AABB b1{v3(0, 0, 0), v3(1, 1, 1)}, b2{v3(0, 2, 0), v3(1, 1, 1)}, b3{v3(0, 3, 0), v3(1, 1, 1)};
    AABB comb = aabb_combine(b1, b2);
    CHECK(overlaps(b1, b2)); CHECK(!overlaps(b1, b3)); CHECK(!contains(b1, b2));
    CHECK(contains(comb, b1)); CHECK(contains(comb, b2)); CHECK(!contains(comb, b3));
}

static void test_geom() {  // geom.rs:1154-1173
    Triangle tri{v3(2, 3.5f, 0), v3(-2, -1.5f, 0), v3(2, -1.5f, 0)};
    CHECK(magnitude2(closest_point(tri, v3(0, 0, 0))) < E);
    Capsule cap{v3(2, 0, 0), v3(4, 0, 0) - v3(2, 0, 0), 1.0f};
    EQV(support(cap, v3(0, 1, 0)), 5.0f, 1.0f, 0.0f);
    EQV(support(cap, v3(-1, 0, 0)), 1.0f, 0.0f, 0.0f);
}

static void test_tensors() {  // physics.rs:321-335
    Mat3 t = tensor(Sphere{v3(0, 0, 0), 1.0f}, 1.0f);
    EQV(t.c[0], 0.4f, 0.0f, 0.0f); EQV(t.c[1], 0.0f, 0.4f, 0.0f); EQV(t.c[2], 0.0f, 0.0f, 0.4f);
}

static void test_compound_rotated_sphere() {  // compound.rs:362-377 restated without Compound:
    // the rotated component is Sphere{(-5,0,0)} / {(5,0,0)} rotate_about(rot, origin); the hit is
    // against the one rotated to (0,5,0).  Pins Quaternion rotate_vector at 1e-6.
    Quat rot = qnormalize(from_arc(v3(1, 0, 0), v3(0, 1, 0)));
    // Volumetric::rotate_about (geom.rs:933-937): set_pos(p + rot*(center - p)) => c += (p' - c)
    Vec3 c0 = v3(5, 0, 0), p = v3(0, 0, 0);
    Vec3 pp = p + rotate_vector(rot, c0 - p);
    Sphere comp{c0 + (pp - c0), 1.0f};
    Moving<Sphere> test_sphere{Sphere{v3(0, 8, 0), 1.0f}, v3(0, -1.5f, 0)};
    Contact col;
    // Compound::contacts -> rhs.contacts(&shape, |c| callback(-c)) with rhs = Moving<Sphere> (:1368)
    bool any = false;
    moving_contacts_static(test_sphere, comp, [&](const Contact& k) { col = neg(k); any = true; });
    CHECK(any);
    REL(col.t, 0.6666663f, E);
    RELV(col.a, 0.0f, 6.0f, 0.0f, E);
}

static void test_compound() {  // compound.rs:362-388, through the restated Compound itself
    std::vector<Component> comps = {Component::sphere(Sphere{v3(-5, 0, 0), 1.0f}), Component::sphere(Sphere{v3(5, 0, 0), 1.0f})};
    Compound compound(comps);
    Moving<Sphere> test_sphere{Sphere{v3(0, 8, 0), 1.0f}, v3(0, -1.5f, 0)};
    CHECK(!compound.contacts(test_sphere, [&](const Contact&) { CHECK(false); }));
    compound.rot = qnormalize(from_arc(v3(1, 0, 0), v3(0, 1, 0)));
    Contact last{}; bool any = compound.contacts(test_sphere, [&](const Contact& c) { last = c; });   // last_contact (collision.rs:477)
    CHECK(any);
    REL(last.t, 0.6666663f, E);
    RELV(last.a, 0.0f, 6.0f, 0.0f, E);
    Rectangle static_rect{v3(0, -2, 0), {v3(1, 0, 0), v3(0, 0, 1)}, {6.0f, 6.0f}};
    compound.rot = quat_one();
    CHECK(compound.contacts(Moving<Rectangle>{static_rect, v3(0, 3, 0)}, [&](const Contact&) {}));   // .unwrap() must not panic
}

int main() {
    test_ray_intersections();
    test_sphere_penetration();
    test_moving_spheres_collision();
    test_sphere_rect_collision();
    test_sphere_tri_collision();
    test_obb_collision();
    test_capsule_moving_sphere();
    test_moving_capsule_collision();
    test_capsule_rect_collision();
    test_capsule_tri_collision();
    test_bvh();
    test_aabb();
    test_geom();
    test_tensors();
    test_compound_rotated_sphere();
    test_compound();
    std::printf("KAT: %d passed, %d failed\n", g_pass, g_fail);
    return g_fail ? 1 : 0;
}
