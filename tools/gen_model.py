#!/usr/bin/env python3
"""Compile the Cassie MJCF (reference: cassie/cassiemujoco/cassie.xml) into a flat C header.

Run at development time only (the reference tree is not present on the GPU box):

    python tools/gen_model.py /root/reference/cassie/cassiemujoco/cassie.xml \
        apex_b200/csrc/cassie_model.h oracle/cassie_model.h

The header holds nothing but numbers: the kinematic tree, inertias, joint/actuator/sensor
parameters, collision primitives and `connect` anchors, in the body / dof / qpos order
MuJoCo would assign (depth-first over the XML), cf. include/cassiemujoco.h:86-158.
The same text is written to the product (`apex_b200/csrc`) and to the oracle (`oracle/`);
the two trees never include each other's files.

Array storage class is left to the includer through CM_ARRAY (C: `static const`,
CUDA: `static __device__ const`).
"""
import re
import sys
import math
import xml.etree.ElementTree as ET
import numpy as np

DEG = math.pi / 180.0


def vec(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None:
        assert len(v) == n, (s, n)
    return v


def quat_from_mat(R):
    # R columns are the frame axes. Robust conversion (Shepperd).
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    q /= np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    return q


def quat_from_xyaxes(v):
    x = v[:3] / np.linalg.norm(v[:3])
    y = v[3:] - x * np.dot(x, v[3:])
    y /= np.linalg.norm(y)
    z = np.cross(x, y)
    return quat_from_mat(np.stack([x, y, z], axis=1))


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat_rot(q, v):
    w, x, y, z = q
    R = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return R @ v


class Model:
    pass


def compile_model(path):
    root = ET.parse(path).getroot()
    comp = root.find('compiler')
    assert comp.get('angle') == 'degree'
    opt = root.find('option')
    m = Model()
    m.timestep = float(opt.get('timestep'))
    m.iterations = int(opt.get('iterations'))
    m.gravity = vec(opt.get('gravity'), 3)

    # defaults that matter (cassie.xml:14-35)
    dflt = root.find('default')
    jnt_limited_default = dflt.find('joint').get('limited') == 'true'
    geom_solref = vec(dflt.find('geom').get('solref'), 2)
    eq_solref = vec(dflt.find('equality').get('solref'), 2)
    m.geom_solref = geom_solref
    m.eq_solref = eq_solref

    bodies = [dict(name='world', parent=0, pos=np.zeros(3), quat=np.array([1., 0, 0, 0]), ipos=np.zeros(3),
                   mass=0.0, inertia=np.zeros(6), dofadr=-1, dofnum=0)]
    joints, dofs, geoms = [], [], []
    qpos0 = []
    sites = {}

    def walk(elem, parent_id):
        for b in elem.findall('body'):
            bid = len(bodies)
            pos = vec(b.get('pos', '0 0 0'), 3)
            if b.get('xyaxes'):
                quat = quat_from_xyaxes(vec(b.get('xyaxes'), 6))
            elif b.get('quat'):
                quat = vec(b.get('quat'), 4)
            else:
                quat = np.array([1., 0, 0, 0])
            ine = b.find('inertial')
            fi = vec(ine.get('fullinertia'), 6)  # Ixx Iyy Izz Ixy Ixz Iyz
            bd = dict(name=b.get('name'), parent=parent_id, pos=pos, quat=quat, ipos=vec(ine.get('pos'), 3),
                      mass=float(ine.get('mass')), inertia=fi, dofadr=len(dofs), dofnum=0)
            bodies.append(bd)
            for j in b.findall('joint'):
                jt = j.get('type', 'hinge')
                jid = len(joints)
                limited = j.get('limited')
                limited = jnt_limited_default if limited is None else (limited == 'true')
                rng = vec(j.get('range', '0 0'), 2)
                ref = float(j.get('ref', '0'))
                if jt in ('hinge', 'ball'):
                    rng = rng * DEG
                    ref = ref * DEG
                jd = dict(name=j.get('name', ''), type={'slide': 0, 'hinge': 1, 'ball': 2}[jt], body=bid,
                          qposadr=len(qpos0), dofadr=len(dofs), axis=vec(j.get('axis', '0 0 1'), 3),
                          pos=vec(j.get('pos', '0 0 0'), 3), ref=ref, limited=limited, range=rng,
                          stiffness=float(j.get('stiffness', '0')), damping=float(j.get('damping', '0')),
                          armature=float(j.get('armature', '0')))
                assert np.all(jd['pos'] == 0)
                joints.append(jd)
                nd = 3 if jt == 'ball' else 1
                if jt == 'ball':
                    qpos0.extend([1, 0, 0, 0])
                else:
                    qpos0.append(ref)
                for k in range(nd):
                    # dof parent: previous dof of this body, else last dof of the nearest ancestor with dofs
                    if bd['dofnum'] > 0:
                        par = len(dofs) - 1
                    else:
                        p = parent_id
                        while p > 0 and bodies[p]['dofnum'] == 0:
                            p = bodies[p]['parent']
                        par = bodies[p]['dofadr'] + bodies[p]['dofnum'] - 1 if p > 0 else -1
                    dofs.append(dict(body=bid, jnt=jid, parent=par, damping=jd['damping'], armature=jd['armature']))
                    bd['dofnum'] += 1
            for g in b.findall('geom'):
                cls = g.get('class', '')
                if not cls.startswith('collision'):
                    continue
                grp = {'collision': 0, 'collision-left': 1, 'collision-right': 2}[cls]
                if g.get('type') == 'sphere':
                    geoms.append(dict(type=0, body=bid, pos=vec(g.get('pos'), 3), axis=np.array([0., 0, 1]),
                                      halflen=0.0, radius=float(g.get('size')), group=grp))
                else:
                    ft = vec(g.get('fromto'), 6)
                    a, c = ft[:3], ft[3:]
                    ax = c - a
                    L = np.linalg.norm(ax)
                    geoms.append(dict(type=1, body=bid, pos=0.5 * (a + c), axis=ax / L, halflen=0.5 * L,
                                      radius=float(g.get('size')), group=grp))
            for s in b.findall('site'):
                sites[s.get('name')] = (bid, vec(s.get('pos'), 3))
            walk(b, bid)

    wb = root.find('worldbody')
    walk(wb, 0)
    floor = wb.find('geom')
    m.floor_pos = vec(floor.get('pos'), 3)
    m.bodies, m.joints, m.dofs, m.geoms = bodies, joints, dofs, geoms
    m.qpos0 = np.array(qpos0, dtype=np.float64)
    m.nbody, m.nq, m.nv, m.njnt = len(bodies), len(qpos0), len(dofs), len(joints)
    assert (m.nbody, m.nq, m.nv) == (26, 35, 32), (m.nbody, m.nq, m.nv)
    m.imu_body, m.imu_pos = sites['imu']

    # world poses at qpos0 (joint displacement zero) for connect anchors (MuJoCo stores the anchor in both bodies)
    xpos = [np.zeros(3)]
    xquat = [np.array([1., 0, 0, 0])]
    for b in bodies[1:]:
        p = b['parent']
        xpos.append(xpos[p] + quat_rot(xquat[p], b['pos']))
        xquat.append(quat_mul(xquat[p], b['quat']))
    name2body = {b['name']: i for i, b in enumerate(bodies)}
    eqs = []
    for c in root.find('equality').findall('connect'):
        b1, b2 = name2body[c.get('body1')], name2body[c.get('body2')]
        a1 = vec(c.get('anchor'), 3)
        w = xpos[b1] + quat_rot(xquat[b1], a1)
        qi = xquat[b2] * np.array([1, -1, -1, -1])
        a2 = quat_rot(qi, w - xpos[b2])
        eqs.append(dict(body1=b1, body2=b2, anchor1=a1, anchor2=a2))
    m.eqs = eqs

    name2jnt = {j['name']: i for i, j in enumerate(joints)}
    acts = []
    for a in root.find('actuator').findall('motor'):
        j = joints[name2jnt[a.get('joint')]]
        cr = vec(a.get('ctrlrange'), 2)
        acts.append(dict(dof=j['dofadr'], qposadr=j['qposadr'], gear=float(a.get('gear')), ctrlmax=cr[1],
                         rpm=float(a.get('user'))))
    m.acts = acts
    act_name = {a.get('name'): i for i, a in enumerate(root.find('actuator').findall('motor'))}
    sens = []
    for s in root.find('sensor'):
        if s.tag == 'actuatorpos':
            sens.append(dict(kind=0, idx=act_name[s.get('actuator')], bits=int(s.get('user'))))
        elif s.tag == 'jointpos':
            j = joints[name2jnt[s.get('joint')]]
            sens.append(dict(kind=1, qposadr=j['qposadr'], dof=j['dofadr'], bits=int(s.get('user'))))
    m.sens = sens
    return m


def fmt(x):
    return repr(float(x))


def arr1(name, vals, ty='double'):
    body = ', '.join((str(int(v)) + ('u' if ty == 'unsigned' else '')) if ty in ('int', 'unsigned') else fmt(v) for v in vals)
    return f'CM_ARRAY {ty} {name}[{len(vals)}] = {{{body}}};\n'


def arr2(name, rows, ty='double'):
    n = len(rows[0])
    body = ',\n  '.join('{' + ', '.join(str(int(v)) if ty == 'int' else fmt(v) for v in r) + '}' for r in rows)
    return f'CM_ARRAY {ty} {name}[{len(rows)}][{n}] = {{\n  {body}}};\n'


def emit(m):
    o = []
    o.append('/* GENERATED by tools/gen_model.py from the Cassie MJCF (reference cassie/cassiemujoco/cassie.xml).\n'
             ' * Numbers only; body/dof/qpos order is MuJoCo\'s depth-first order (include/cassiemujoco.h:86-158).\n'
             ' * Do not edit by hand. */\n')
    o.append('#ifndef CASSIE_MODEL_H\n#define CASSIE_MODEL_H\n#ifndef CM_ARRAY\n#define CM_ARRAY static const\n#endif\n')
    o.append(f'#define CM_NBODY {m.nbody}\n#define CM_NQ {m.nq}\n#define CM_NV {m.nv}\n#define CM_NJNT {m.njnt}\n'
             f'#define CM_NU {len(m.acts)}\n#define CM_NGEOM {len(m.geoms)}\n#define CM_NEQ {len(m.eqs)}\n'
             f'#define CM_TIMESTEP {fmt(m.timestep)}\n#define CM_ITERATIONS {m.iterations}\n'
             f'#define CM_GRAVITY_Z {fmt(m.gravity[2])}\n'
             f'#define CM_FLOOR_Z {fmt(m.floor_pos[2])}\n'
             f'#define CM_GEOM_SOLREF_TC {fmt(m.geom_solref[0])}\n#define CM_GEOM_SOLREF_DR {fmt(m.geom_solref[1])}\n'
             f'#define CM_EQ_SOLREF_TC {fmt(m.eq_solref[0])}\n#define CM_EQ_SOLREF_DR {fmt(m.eq_solref[1])}\n'
             '#define CM_LIMIT_SOLREF_TC 0.02\n#define CM_LIMIT_SOLREF_DR 1.0\n'
             '#define CM_SOLIMP_DMIN 0.9\n#define CM_SOLIMP_DMAX 0.95\n#define CM_SOLIMP_WIDTH 0.001\n'
             '#define CM_SOLIMP_MID 0.5\n#define CM_SOLIMP_POWER 2.0\n'
             f'#define CM_IMU_BODY {m.imu_body}\n')
    o.append(arr1('CM_imu_pos', m.imu_pos))
    B = m.bodies
    o.append(arr1('CM_body_parent', [b['parent'] for b in B], 'int'))
    o.append(arr1('CM_body_dofadr', [b['dofadr'] for b in B], 'int'))
    o.append(arr1('CM_body_dofnum', [b['dofnum'] for b in B], 'int'))
    o.append(arr2('CM_body_pos', [b['pos'] for b in B]))
    o.append(arr2('CM_body_quat', [b['quat'] for b in B]))
    o.append(arr2('CM_body_ipos', [b['ipos'] for b in B]))
    o.append(arr1('CM_body_mass', [b['mass'] for b in B]))
    o.append('/* Ixx Iyy Izz Ixy Ixz Iyz about the body com, in body-frame axes */\n')
    o.append(arr2('CM_body_inertia', [b['inertia'] for b in B]))
    J = m.joints
    o.append('/* joint type: 0 slide, 1 hinge, 2 ball */\n')
    o.append(arr1('CM_jnt_type', [j['type'] for j in J], 'int'))
    o.append(arr1('CM_jnt_body', [j['body'] for j in J], 'int'))
    o.append(arr1('CM_jnt_qposadr', [j['qposadr'] for j in J], 'int'))
    o.append(arr1('CM_jnt_dofadr', [j['dofadr'] for j in J], 'int'))
    o.append(arr2('CM_jnt_axis', [j['axis'] for j in J]))
    o.append(arr1('CM_jnt_limited', [int(j['limited']) for j in J], 'int'))
    o.append(arr2('CM_jnt_range', [j['range'] for j in J]))
    o.append(arr1('CM_jnt_stiffness', [j['stiffness'] for j in J]))
    D = m.dofs
    o.append(arr1('CM_dof_body', [d['body'] for d in D], 'int'))
    o.append(arr1('CM_dof_jnt', [d['jnt'] for d in D], 'int'))
    o.append(arr1('CM_dof_parent', [d['parent'] for d in D], 'int'))
    o.append(arr1('CM_dof_damping', [d['damping'] for d in D]))
    o.append(arr1('CM_dof_armature', [d['armature'] for d in D]))
    o.append(arr1('CM_qpos0', m.qpos0))
    o.append('/* fixed start pose written by cassie_sim_set_const (libcassiemujoco.so .rodata @0x2ed40) */\n')
    o.append(arr1('CM_qpos_init', INIT_QPOS))
    A = m.acts
    o.append(arr1('CM_act_dof', [a['dof'] for a in A], 'int'))
    o.append(arr1('CM_act_qposadr', [a['qposadr'] for a in A], 'int'))
    o.append(arr1('CM_act_gear', [a['gear'] for a in A]))
    o.append(arr1('CM_act_ctrlmax', [a['ctrlmax'] for a in A]))
    o.append(arr1('CM_act_rpm', [a['rpm'] for a in A]))
    drive = [s for s in m.sens if s['kind'] == 0]
    joint = [s for s in m.sens if s['kind'] == 1]
    assert [s['idx'] for s in drive] == list(range(10))
    o.append(arr1('CM_drive_bits', [s['bits'] for s in drive], 'int'))
    o.append(arr1('CM_jsens_qposadr', [s['qposadr'] for s in joint], 'int'))
    o.append(arr1('CM_jsens_dof', [s['dof'] for s in joint], 'int'))
    o.append(arr1('CM_jsens_bits', [s['bits'] for s in joint], 'int'))
    G = m.geoms
    o.append('/* collision primitives: type 0 sphere / 1 capsule; group 0 = floor only, 1 = left leg, 2 = right leg */\n')
    o.append(arr1('CM_geom_type', [g['type'] for g in G], 'int'))
    o.append(arr1('CM_geom_body', [g['body'] for g in G], 'int'))
    o.append(arr1('CM_geom_group', [g['group'] for g in G], 'int'))
    o.append(arr2('CM_geom_pos', [g['pos'] for g in G]))
    o.append(arr2('CM_geom_axis', [g['axis'] for g in G]))
    o.append(arr1('CM_geom_halflen', [g['halflen'] for g in G]))
    o.append(arr1('CM_geom_radius', [g['radius'] for g in G]))
    E = m.eqs
    o.append(arr1('CM_eq_body1', [e['body1'] for e in E], 'int'))
    o.append(arr1('CM_eq_body2', [e['body2'] for e in E], 'int'))
    o.append(arr2('CM_eq_anchor1', [e['anchor1'] for e in E]))
    o.append(arr2('CM_eq_anchor2', [e['anchor2'] for e in E]))

    # ---- derived tables for the warp-per-env kernel ----
    depth = [0] * m.nbody
    for i, b in enumerate(B):
        if i > 0:
            depth[i] = depth[b['parent']] + 1
    o.append(arr1('CM_body_level', depth, 'int'))
    o.append(f'#define CM_MAXLEVEL {max(depth)}\n')
    # joint owned by each body (pelvis is special-cased: 3 slides + ball); -1 when the body has no joint
    bj = [-1] * m.nbody
    for j, jn in enumerate(J):
        bj[jn['body']] = j
    o.append(arr1('CM_body_jnt', bj, 'int'))
    # ancestor lists of every dof (root first, excluding the dof itself)
    anc = []
    for i in range(m.nv):
        ch = []
        p = D[i]['parent']
        while p >= 0:
            ch.append(p)
            p = D[p]['parent']
        anc.append(ch[::-1])
    maxd = max(len(a) for a in anc)
    o.append(f'#define CM_MAXANC {maxd}\n')
    o.append(arr1('CM_dof_nanc', [len(a) for a in anc], 'int'))
    o.append(arr2('CM_dof_anc', [a + [-1] * (maxd - len(a)) for a in anc], 'int'))
    rp = row_pointers(anc)
    o.append(f'#define CM_MNNZ {rp[-1]}\n')
    o.append(f'#define CM_LEG_ROWSPAN {rp[19] - rp[6]} /* CM_dof_rowptr[19 + s] - CM_dof_rowptr[6 + s]: the two legs have the same row layout */\n')
    assert all(rp[19 + s_] - rp[6 + s_] == rp[19] - rp[6] for s_ in range(13))
    o.append(arr1('CM_dof_rowptr', rp, 'int'))
    o.append(arr1('CM_dof_ancmask', [sum(1 << k for k in a) for a in anc], 'unsigned'))
    # last dof on the path to each body (bodies without dofs inherit their parent's)
    last = []
    for i, b in enumerate(B):
        k = i
        while k > 0 and B[k]['dofnum'] == 0:
            k = B[k]['parent']
        last.append(B[k]['dofadr'] + B[k]['dofnum'] - 1 if k > 0 else -1)
    o.append(arr1('CM_body_lastdof', last, 'int'))
    o.append(arr1('CM_body_dofmask', [0 if l < 0 else ((sum(1 << k for k in anc[l])) | (1 << l)) for l in last], 'unsigned'))
    kids = [[] for _ in B]
    for i, b in enumerate(B):
        if i > 0:
            kids[b['parent']].append(i)
    mk = max(len(k) for k in kids[1:])
    o.append(f'#define CM_MAXCHILD {mk}\n')
    o.append(arr1('CM_body_nchild', [0] + [len(k) for k in kids[1:]], 'int'))
    o.append(arr2('CM_body_child', [[-1] * mk] + [k + [-1] * (mk - len(k)) for k in kids[1:]], 'int'))
    # pointer-jumping table for path sums down the body tree: byte r of entry b = the 2^r-th ancestor of body b, 31 where the
    # path is shorter (lane 31 is no body and carries zeros); 4 rounds cover the 10 bodies of the longest path
    def nth_parent(b, n_):
        for _ in range(n_):
            if b <= 0:
                return 31
            b = B[b]['parent']
        return b
    jump = []
    for b_ in range(32):
        if b_ >= len(B):
            jump.append(31 | 31 << 8 | 31 << 16 | 31 << 24)
        else:
            jump.append(sum((nth_parent(b_, 1 << r_) & 0xff) << (8 * r_) for r_ in range(4)))
    o.append(arr1('CM_body_jump', jump, 'unsigned'))
    # bodies are numbered depth-first, so a body's subtree is the contiguous range [b, b + CM_body_subtree[b])
    sub = [1] * len(B)
    for b_ in range(len(B) - 1, 0, -1):
        sub[B[b_]['parent']] += sub[b_]
    for b_ in range(1, len(B)):
        assert all(B[c_]['parent'] >= b_ for c_ in range(b_ + 1, b_ + sub[b_])) and (b_ + sub[b_] == len(B) or B[b_ + sub[b_]]['parent'] < b_)
    o.append(arr1('CM_body_subtree', sub + [0] * (32 - len(B)), 'int'))
    ddepth = [len(a) for a in anc]
    o.append(arr1('CM_dof_armature_f', [d['armature'] for d in D]))
    o.append(arr1('CM_dof_qposadr', [J[d['jnt']]['qposadr'] + (i - J[d['jnt']]['dofadr'] if J[d['jnt']]['type'] != 2 else 0) for i, d in enumerate(D)], 'int'))
    o.append('#endif\n')
    return ''.join(o)


def row_pointers(anc):
    """Start of every dof's row in the tree-sparse storage (nv + 1 entries; the last one is the total size).  The six base
    rows are packed; every leg row starts on a multiple of four words so that a row can be moved with 16-byte accesses."""
    rp = [0]
    for i, a in enumerate(anc):
        end = rp[-1] + len(a)
        if i + 1 >= 6:
            end = (end + 3) // 4 * 4
        rp.append(end)
    return rp


def emit_gen(m):
    """Straight-line code for the sparse L^T D L factorisation and the row half-solve (indices are compile-time)."""
    D = m.dofs
    nv = m.nv
    anc = []
    for i in range(nv):
        ch = []
        p = D[i]['parent']
        while p >= 0:
            ch.append(p)
            p = D[p]['parent']
        anc.append(ch[::-1])
    o = ['/* GENERATED by tools/gen_model.py — do not edit.  Unrolled tree-sparse kernels for the Cassie dof tree.\n'
         ' * Storage convention: w.Ms[CM_dof_rowptr[k] + t] holds U = D_k * L[k][j] for the t-th ancestor j of k (root first);\n * w.Dinv[k] = 1/D_k. */\n'
         '#ifndef CASSIE_GEN_H\n#define CASSIE_GEN_H\n']
    # ---- leg-ancestor masks for the two-legs-at-once factorisation (bit = left-leg dof index) ----
    masks = []
    for s_ in range(13):
        kL, kR = 6 + s_, 19 + s_
        assert [a + 13 if a >= 6 else a for a in anc[kL]] == anc[kR]
        masks.append(sum(1 << a for a in anc[kL] if a >= 6))
    o.append('CM_ARRAY unsigned CM_leg_ancmask[13] = {' + ', '.join(f'{x}u' for x in masks) + '};\n\n')
    # ---- half solve: y <- L^-T y with y in registers ----
    rowptr = row_pointers(anc)
    o.append('/* y <- L^-T y for one row held in registers (all indices compile-time) */\n')
    o.append('template <typename T> CW_FN void cw_half_solve_regs(const CassieWs<T> &w, T *y) {\n')
    for k in range(nv - 1, 0, -1):
        o.append(f'  {{ const T s = y[{k}] * w.Dinv[{k}];')
        for t, j in enumerate(anc[k]):
            o.append(f' y[{j}] -= w.Ms[{rowptr[k] + t}] * s;')
        o.append(' }\n')
    o.append('}\n#endif\n')
    return ''.join(o)


def emit_f32(text):
    """float copies (NAME_f32) of every double table of cassie_model.h, for the float32 kernel: the device code would otherwise
    load 8-byte constants and convert them (F2F.F32.F64 runs on the slow fp64 pipe) at every use."""
    o = ['/* GENERATED by tools/gen_model.py from cassie_model.h: float32 copies of its double tables.  Do not edit. */\n'
         '#ifndef CASSIE_MODEL_F32_H\n#define CASSIE_MODEL_F32_H\n']
    for m in re.finditer(r'CM_ARRAY double (\w+)((?:\[\d+\])+) = (\{.*?\});', text, re.S):
        body = re.sub(r'(?<![\w.])(-?\d+\.\d*(?:e[-+]?\d+)?|-?\d+e[-+]?\d+)(?![\w.])', lambda k: k.group(1) + 'f', m.group(3))
        o.append(f'CM_ARRAY float {m.group(1)}_f32{m.group(2)} = {body};\n')
    o.append('#endif\n')
    return ''.join(o)


INIT_QPOS = [0.0, 0.0, 1.01, 1.0, 0.0, 0.0, 0.0, 0.0045, 0.0, 0.4973, 0.9785, -0.0164, 0.01787, -0.2049, -1.1997, 0.0,
             1.4267, 0.0, -1.5244, 1.5244, -1.5968, -0.0045, 0.0, 0.4973, 0.9786, 0.00386, -0.01524, -0.2051, -1.1997,
             0.0, 1.4267, 0.0, -1.5244, 1.5244, -1.5968]

if __name__ == '__main__':
    if sys.argv[1] == '--f32':  # python tools/gen_model.py --f32 apex_b200/csrc/cassie_model.h apex_b200/csrc/cassie_model_f32.h
        with open(sys.argv[3], 'w') as f:
            f.write(emit_f32(open(sys.argv[2]).read()))
        sys.exit(0)
    mdl = compile_model(sys.argv[1])
    text = emit(mdl)
    for out in sys.argv[2:]:
        with open(out, 'w') as f:
            f.write(text)
    import os
    with open(os.path.join(os.path.dirname(sys.argv[2]), 'cassie_gen.h'), 'w') as f:
        f.write(emit_gen(mdl))
    print(f'nbody={mdl.nbody} nq={mdl.nq} nv={mdl.nv} njnt={mdl.njnt} ngeom={len(mdl.geoms)}')
