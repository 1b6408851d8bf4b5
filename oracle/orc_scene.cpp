/* TEST INFRASTRUCTURE — CPU oracle: asset blob loader (format defined in real_robots_b200/blob.py). */
#include "orc_scene.h"
#include <cstdio>
#include <cstring>

bool OrcBlob::load(const char* path, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    raw.resize((size_t)n);
    if (std::fread(raw.data(), 1, (size_t)n, f) != (size_t)n) { std::fclose(f); err = "short read"; return false; }
    std::fclose(f);
    if (n < 16 || std::memcmp(raw.data(), "RRB200\0\0", 8) != 0) { err = "bad magic"; return false; }
    uint32_t ver, cnt;
    std::memcpy(&ver, raw.data() + 8, 4);
    std::memcpy(&cnt, raw.data() + 12, 4);
    if (ver != 1) { err = "bad version"; return false; }
    entries.clear();
    for (uint32_t i = 0; i < cnt; i++) {
        const uint8_t* e = raw.data() + 16 + 112 * (size_t)i;
        OrcBlobEntry en;
        char nm[49]; std::memcpy(nm, e, 48); nm[48] = 0;
        en.name = nm;
        uint32_t dt, nd; std::memcpy(&dt, e + 48, 4); std::memcpy(&nd, e + 52, 4);
        en.dtype = (int)dt; en.ndim = (int)nd;
        std::memcpy(en.shape, e + 56, 32);
        uint64_t off; std::memcpy(&off, e + 88, 8); std::memcpy(&en.nbytes, e + 96, 8);
        if (off + en.nbytes > (uint64_t)n) { err = "entry out of range"; return false; }
        en.data = raw.data() + off;
        entries.push_back(en);
    }
    return true;
}
const OrcBlobEntry* OrcBlob::find(const char* name) const {
    for (size_t i = 0; i < entries.size(); i++) if (entries[i].name == name) return &entries[i];
    return nullptr;
}
static const uint8_t* need(const OrcBlob* b, const char* name, int dtype) {
    const OrcBlobEntry* e = b->find(name);
    if (!e || e->dtype != dtype) { std::fprintf(stderr, "oracle: blob entry %s missing/wrong type\n", name); std::abort(); }
    return e->data;
}
const double* OrcBlob::f64(const char* n) const { return (const double*)need(this, n, 3); }
const float* OrcBlob::f32(const char* n) const { return (const float*)need(this, n, 2); }
const int32_t* OrcBlob::i32(const char* n) const { return (const int32_t*)need(this, n, 1); }
const int64_t* OrcBlob::i64(const char* n) const { return (const int64_t*)need(this, n, 4); }
const uint8_t* OrcBlob::u8(const char* n) const { return need(this, n, 0); }
uint64_t OrcBlob::dim(const char* n, int i) const { const OrcBlobEntry* e = find(n); return e ? e->shape[i] : 0; }

void orc_default_config(OrcConfig* c) {
    c->dt = 0.005;
    c->solver_iterations = 50;
    c->erp = 0.2;
    c->erp2 = 0.08;
    c->linear_slop = 1e-5;
    c->warmstart = 0.1;
    c->residual_threshold = 1e-7;
    c->lin_damping = 0.04;
    c->ang_damping = 0.04;
    c->max_coord_vel = 100.0;
    c->motor_kp = 0.1;
    c->motor_kd = 1.0;
    c->motor_max_force = 100000.0;
    c->limit_max_impulse = 100.0;
    c->use_box_inertia = 1;
    c->gravity = -9.81;
    c->enable_contacts = 1;
    c->contact_breaking = 0.02;
    c->margin = 0.001;
    c->restitution_threshold = 0.2;
    c->friction_erp = 0.2;
    c->torsional_friction = 1;
}

static V3 v3at(const double* p, int i) { return V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

bool OrcScene::load(const char* path, int nobj, const OrcConfig& cfg, std::string& err) {
    if (nobj < 1 || nobj > ORC_MAXOBJ) { err = "n_objects must be 1..3"; return false; }
    if (!blob.load(path, err)) return false;
    n_objects = nobj;
    if (blob.dim("robot.parent", 0) != ORC_NL) { err = "robot link count"; return false; }
    const int32_t* par = blob.i32("robot.parent");
    const int32_t* jt = blob.i32("robot.jtype");
    const int32_t* df = blob.i32("robot.dof");
    const double* r0 = blob.f64("robot.rot0");
    const double* inr = blob.f64(cfg.use_box_inertia ? "robot.inertia_box" : "robot.inertia_urdf");
    for (int i = 0; i < ORC_NL; i++) {
        parent[i] = par[i]; jtype[i] = jt[i]; dof[i] = df[i];
        axis[i] = v3at(blob.f64("robot.axis"), i);
        dvec[i] = v3at(blob.f64("robot.dvec"), i);
        evec[i] = v3at(blob.f64("robot.evec"), i);
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) rot0[i].m[a][b] = r0[9 * i + 3 * a + b];
        mass[i] = blob.f64("robot.mass")[i];
        inertia[i] = v3at(inr, i);
        damping[i] = blob.f64("robot.damping")[i];
        lower[i] = blob.f64("robot.lower")[i];
        upper[i] = blob.f64("robot.upper")[i];
        if (dof[i] >= 0) link_of_dof[dof[i]] = i;
    }
    base_pos = v3at(blob.f64("robot.base_pos"), 0);
    for (int i = 0; i <= ORC_NL; i++) shape_of_link[i] = blob.i32("robot.shape_of_link")[i];
    table_pos = v3at(blob.f64("table.pos"), 0);
    const double* tq = blob.f64("table.quat");
    table_quat = Q4(tq[0], tq[1], tq[2], tq[3]);
    table_upper_offset = v3at(blob.f64("table.upper_offset"), 0);
    table_shape[0] = blob.i32("table.shape")[0];
    table_shape[1] = blob.i32("table.shape")[1];
    for (int o = 0; o < ORC_MAXOBJ; o++) {
        obj_mass[o] = blob.f64("obj.mass")[o];
        obj_inertia[o] = v3at(blob.f64("obj.inertia"), o);
        const double* p = blob.f64("obj.init_pose") + 7 * o;
        obj_init_pos[o] = V3(p[0], p[1], p[2]);
        obj_init_quat[o] = Q4(p[3], p[4], p[5], p[6]);
        obj_shape[o] = blob.i32("obj.shape")[o];
    }
    int S = (int)blob.dim("shape.body", 0);
    shapes.resize(S);
    for (int s = 0; s < S; s++) {
        OrcShape& sh = shapes[s];
        sh.body = blob.i32("shape.body")[s]; sh.link = blob.i32("shape.link")[s];
        sh.tri_start = blob.i32("shape.tri_start")[s]; sh.tri_count = blob.i32("shape.tri_count")[s];
        sh.hull_start = blob.i32("shape.hull_start")[s]; sh.hull_count = blob.i32("shape.hull_count")[s];
        sh.tex = blob.i32("shape.tex")[s];
        for (int k = 0; k < 4; k++) sh.rgba[k] = blob.f32("shape.rgba")[4 * s + k];
        sh.friction = blob.f64("shape.friction")[s]; sh.restitution = blob.f64("shape.restitution")[s];
        sh.rolling = blob.f64("shape.rolling")[s]; sh.spinning = blob.f64("shape.spinning")[s];
    }
    tris = blob.f32("tri.data");
    hull_verts = blob.f64("hull.verts");
    int K = (int)blob.dim("tex.meta", 0);
    textures.resize(K);
    for (int k = 0; k < K; k++) {
        const int64_t* m = blob.i64("tex.meta") + 4 * k;
        textures[k].w = (int)m[0]; textures[k].h = (int)m[1];
        textures[k].data = blob.u8("tex.data") + m[2];
        textures[k].uniform = m[3] != 0;
    }
    return true;
}
