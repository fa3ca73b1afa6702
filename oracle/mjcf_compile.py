"""ORACLE (test infrastructure, not product code): MJCF scene merge + model compile.

Restates, for the RoboSumo scenes only, what the reference does before physics starts:

  * scene merge        -- robosumo/robosumo/envs/utils.py:46-183  (construct_scene)
  * env registration   -- robosumo/robosumo/__init__.py:8-105      (densities, tatami_size)
  * scope naming       -- robosumo/robosumo/envs/sumo.py:65-75
  * MuJoCo 2.1 model compilation of the merged MJCF [M] (defaults, degree->rad,
    fromto capsules, inertiafromgeom, depth-first body/joint/dof/geom ordering).
    MuJoCo itself is not in the reference tree; the rules applied here are the
    published MJCF semantics, restricted to the features these four XML files use.

The output is a plain dict of lists ("generic model") that `physics_oracle.c`
consumes and that is committed as JSON under tests/golden/ (made by
tests/golden/make_model_golden.py, which is the only code that reads
/root/reference XML files).

Only tests/, bench.py's cpu_baseline/reference legs and __graft_entry__.smoke()
may import this module.  Parity status: UNPINNED against MuJoCo (no MuJoCo in the
build container); pinned only against hand-derived mass anchors.
"""
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

# mjtGeom / mjtJoint enum values [M]
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX = 0, 2, 3, 5, 6
JNT_FREE, JNT_HINGE = 0, 3

AGENT_DENSITY = {'ant': 13.0, 'bug': 10.0, 'spider': 39.0}   # robosumo/__init__.py


def _floats(s):
    return [float(x) for x in s.split()]


def _quat_z2vec(vec):
    """Quaternion rotating the z axis onto `vec` (minimal rotation) [M: mjuu_z2quat]."""
    v = np.asarray(vec, dtype=np.float64)
    v = v / np.linalg.norm(v)
    z = np.array([0.0, 0.0, 1.0])
    axis = np.cross(z, v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        if v[2] > 0:
            return [1.0, 0.0, 0.0, 0.0]
        return [0.0, 1.0, 0.0, 0.0]
    axis /= s
    ang = math.atan2(s, v[2])
    return [math.cos(ang / 2)] + list(math.sin(ang / 2) * axis)


def merge_scene(assets_dir, agent_names, tatami_size=2.0):
    """construct_scene (utils.py:46-183) without the colour handling.

    Returns (scene_root, scopes, densities)."""
    scene_root = ET.parse(os.path.join(assets_dir, 'tatami.xml')).getroot()
    scene_default = scene_root.find('default')
    scene_body = scene_root.find('worldbody')
    n = len(agent_names)
    assert n == 2
    # tatami resize (utils.py:64-88): formatted through "%.2f"
    for geom in scene_body.findall('geom'):
        nm = geom.get('name')
        s = float("%.2f" % tatami_size)
        if nm == 'tatami':
            b = float("%.2f" % (tatami_size + 0.3))
            geom.set('size', "%r %r 0.25" % (b, b))
        elif nm == 'topborder':
            geom.set('fromto', "%r %r 0.5 %r %r 0.5" % (-s, s, s, s))
        elif nm == 'rightborder':
            geom.set('fromto', "%r %r 0.5 %r %r 0.5" % (s, -s, s, s))
        elif nm == 'bottomborder':
            geom.set('fromto', "%r %r 0.5 %r %r 0.5" % (-s, -s, s, -s))
        elif nm == 'leftborder':
            geom.set('fromto', "%r %r 0.5 %r %r 0.5" % (-s, -s, -s, s))
    scopes = ["%s%d" % (name, i) for i, name in enumerate(agent_names)]   # sumo.py:65-68
    densities = [AGENT_DENSITY[a] for a in agent_names]
    # init poses (utils.py:108-115)
    r, phi, z = 1.5, 0.0, 0.75
    delta = 2.0 * math.pi / n
    poses = [(r * math.cos(phi + i * delta), r * math.sin(phi + i * delta), z) for i in range(n)]
    actuators = []
    for i in range(n):
        agent = ET.parse(os.path.join(assets_dir, agent_names[i] + '.xml')).getroot()
        cls = ET.SubElement(scene_default, 'default', attrib={'class': scopes[i]})
        for child in list(agent.find('default')):
            if child.tag == 'geom':
                child.set('density', str(densities[i]))
            cls.append(child)
        body = agent.find('body')
        body.set('pos', " ".join(str(x) for x in poses[i]))
        for el in body.iter():
            if el.tag == 'geom':
                el.set('class', scopes[i])
            if el.get('name') is not None:
                el.set('name', scopes[i] + '/' + el.get('name'))
        scene_body.append(body)
        for motor in list(agent.find('actuator')):
            motor.set('joint', scopes[i] + '/' + motor.get('joint'))
            motor.set('class', scopes[i])
            actuators.append(motor)
    return scene_root, scopes, actuators


def compile_model(assets_dir, agent_names, tatami_size=2.0):
    root, scopes, actuators = merge_scene(assets_dir, agent_names, tatami_size)
    comp = root.find('compiler')
    assert comp.get('angle') == 'degree' and comp.get('coordinate') == 'local'
    assert comp.get('inertiafromgeom') == 'true'
    opt = root.find('option')
    assert opt.get('integrator') == 'RK4'
    timestep = float(opt.get('timestep'))
    deg = math.pi / 180.0

    # defaults: top-level <default> then one nested class per agent
    top = root.find('default')
    jdef = dict(armature=0.0, damping=0.0, limited=False, margin=0.0)
    tj = top.find('joint')
    if tj is not None:
        jdef['armature'] = float(tj.get('armature', jdef['armature']))
        jdef['damping'] = float(tj.get('damping', jdef['damping']))
        jdef['limited'] = tj.get('limited', 'false') == 'true'
    gdef_main = dict(contype=1, conaffinity=1, condim=3, density=1000.0,
                     friction=[1.0, 0.005, 0.0001], margin=0.0)
    gdef = {None: gdef_main}
    for cls in top.findall('default'):
        d = dict(gdef_main)
        g = cls.find('geom')
        if g is not None:
            for k in ('contype', 'conaffinity', 'condim'):
                if g.get(k) is not None:
                    d[k] = int(g.get(k))
            if g.get('density') is not None:
                d['density'] = float(g.get('density'))
            if g.get('margin') is not None:
                d['margin'] = float(g.get('margin'))
            if g.get('friction') is not None:
                d['friction'] = _floats(g.get('friction'))
        gdef[cls.get('class')] = d

    M = dict(body_name=[], body_parent=[], body_pos=[], body_quat=[], body_ipos=[], body_iquat=[],
             body_mass=[], body_inertia=[], body_jntadr=[], body_jntnum=[], body_dofadr=[], body_dofnum=[],
             body_weldid=[], body_rootid=[],
             jnt_name=[], jnt_type=[], jnt_qposadr=[], jnt_dofadr=[], jnt_bodyid=[], jnt_pos=[], jnt_axis=[],
             jnt_range=[], jnt_limited=[], jnt_margin=[],
             dof_bodyid=[], dof_jntid=[], dof_armature=[], dof_damping=[],
             geom_name=[], geom_type=[], geom_bodyid=[], geom_pos=[], geom_quat=[], geom_size=[],
             geom_margin=[], geom_friction=[], geom_contype=[], geom_conaffinity=[], geom_condim=[],
             qpos0=[])
    nq = [0]
    nv = [0]
    pending_geoms = []   # (bodyid, elem) : geoms are numbered body by body

    def add_geom(el, bodyid):
        d = gdef[el.get('class')]
        gt = {'plane': GEOM_PLANE, 'sphere': GEOM_SPHERE, 'capsule': GEOM_CAPSULE,
              'cylinder': GEOM_CYLINDER, 'box': GEOM_BOX}[el.get('type', 'sphere')]
        size = _floats(el.get('size'))
        pos = _floats(el.get('pos', '0 0 0'))
        quat = [1.0, 0.0, 0.0, 0.0]
        if el.get('fromto') is not None:
            ft = _floats(el.get('fromto'))
            a, b = np.array(ft[:3]), np.array(ft[3:])
            pos = list((a + b) / 2)
            size = [size[0], float(np.linalg.norm(b - a) / 2), 0.0]
            quat = _quat_z2vec(a - b)            # [M] vec = from - to
        size = (size + [0.0, 0.0, 0.0])[:3]
        density = float(el.get('density')) if el.get('density') is not None else d['density']
        friction = _floats(el.get('friction')) if el.get('friction') is not None else d['friction']
        M['geom_name'].append(el.get('name'))
        M['geom_type'].append(gt)
        M['geom_bodyid'].append(bodyid)
        M['geom_pos'].append(pos)
        M['geom_quat'].append(quat)
        M['geom_size'].append(size)
        M['geom_margin'].append(float(el.get('margin')) if el.get('margin') is not None else d['margin'])
        M['geom_friction'].append(friction)
        M['geom_contype'].append(int(el.get('contype')) if el.get('contype') is not None else d['contype'])
        M['geom_conaffinity'].append(int(el.get('conaffinity')) if el.get('conaffinity') is not None else d['conaffinity'])
        M['geom_condim'].append(int(el.get('condim')) if el.get('condim') is not None else d['condim'])
        # mass / inertia [M: mjCGeom::GetVolume / SetInertia]
        r = size[0]
        if gt == GEOM_SPHERE:
            mass = density * 4.0 / 3.0 * math.pi * r ** 3
            inertia = [0.4 * mass * r * r] * 3
        elif gt == GEOM_CAPSULE:
            h = 2.0 * size[1]
            mass = density * (math.pi * r * r * h + 4.0 / 3.0 * math.pi * r ** 3)
            ms = mass * 4 * r / (4 * r + 3 * h)
            mc = mass - ms
            ixx = mc * (3 * r * r + h * h) / 12 + 2 * ms * r * r / 5 + ms * h * (3 * r + 2 * h) / 8
            izz = mc * r * r / 2 + 2 * ms * r * r / 5
            inertia = [ixx, ixx, izz]
        else:
            mass, inertia = 0.0, [0.0, 0.0, 0.0]   # static world geoms
        return mass, inertia, pos, quat

    def add_body(el, parent, weld_parent, rootid):
        bid = len(M['body_parent'])
        joints = el.findall('joint') if el.tag == 'body' else []
        M['body_name'].append(el.get('name', 'world'))
        M['body_parent'].append(parent)
        M['body_pos'].append(_floats(el.get('pos', '0 0 0')) if el.tag == 'body' else [0.0, 0.0, 0.0])
        M['body_quat'].append([1.0, 0.0, 0.0, 0.0])
        assert el.get('quat') is None and el.get('euler') is None
        M['body_jntadr'].append(len(M['jnt_type']) if joints else -1)
        M['body_jntnum'].append(len(joints))
        M['body_dofadr'].append(nv[0] if joints else -1)
        weld = bid if (joints or bid == 0) else weld_parent
        M['body_weldid'].append(weld)
        if bid == 0:
            rootid = 0
        elif parent == 0:
            rootid = bid
        M['body_rootid'].append(rootid)
        ndof = 0
        for j in joints:
            jt = j.get('type', 'hinge')
            jid = len(M['jnt_type'])
            M['jnt_name'].append(j.get('name'))
            M['jnt_bodyid'].append(bid)
            M['jnt_qposadr'].append(nq[0])
            M['jnt_dofadr'].append(nv[0])
            M['jnt_pos'].append(_floats(j.get('pos', '0 0 0')))
            arm = float(j.get('armature')) if j.get('armature') is not None else jdef['armature']
            dmp = float(j.get('damping')) if j.get('damping') is not None else jdef['damping']
            lim = (j.get('limited') == 'true') if j.get('limited') is not None else jdef['limited']
            M['jnt_margin'].append(float(j.get('margin')) if j.get('margin') is not None else jdef['margin'])
            if jt == 'free':
                M['jnt_type'].append(JNT_FREE)
                M['jnt_axis'].append([0.0, 0.0, 1.0])
                M['jnt_range'].append([0.0, 0.0])
                M['jnt_limited'].append(0)
                M['qpos0'] += M['body_pos'][bid] + [1.0, 0.0, 0.0, 0.0]
                nq[0] += 7
                k = 6
            else:
                assert jt == 'hinge'
                ax = np.array(_floats(j.get('axis')))
                M['jnt_type'].append(JNT_HINGE)
                M['jnt_axis'].append(list(ax / np.linalg.norm(ax)))
                rg = _floats(j.get('range', '0 0'))
                M['jnt_range'].append([rg[0] * deg, rg[1] * deg])
                M['jnt_limited'].append(1 if lim else 0)
                M['qpos0'].append(0.0)
                nq[0] += 1
                k = 1
            for _ in range(k):
                M['dof_bodyid'].append(bid)
                M['dof_jntid'].append(jid)
                M['dof_armature'].append(arm)
                M['dof_damping'].append(dmp)
            nv[0] += k
            ndof += k
        M['body_dofnum'].append(ndof)
        geoms = el.findall('geom')
        for g in geoms:
            pending_geoms.append((bid, g))
        M['body_ipos'].append(None)
        M['body_iquat'].append(None)
        M['body_mass'].append(None)
        M['body_inertia'].append(None)
        for child in el.findall('body'):
            add_body(child, bid, weld, rootid)
        return bid

    add_body(root.find('worldbody'), 0, 0, 0)
    # geoms are stored grouped by body id (MuJoCo numbers geoms in body order)
    per_body = {}
    for bid, g in pending_geoms:
        per_body.setdefault(bid, []).append(g)
    for bid in range(len(M['body_parent'])):
        gl = per_body.get(bid, [])
        res = [add_geom(g, bid) for g in gl]
        if bid == 0:
            M['body_mass'][0] = 0.0
            M['body_inertia'][0] = [0.0, 0.0, 0.0]
            M['body_ipos'][0] = [0.0, 0.0, 0.0]
            M['body_iquat'][0] = [1.0, 0.0, 0.0, 0.0]
            continue
        assert len(res) == 1, "RoboSumo bodies carry exactly one geom"
        mass, inertia, pos, quat = res[0]
        M['body_mass'][bid] = mass
        M['body_inertia'][bid] = inertia
        M['body_ipos'][bid] = pos
        M['body_iquat'][bid] = quat

    # actuators (motors): joint transmission, gear, ctrlrange
    M['act_jntid'] = []
    M['act_gear'] = []
    M['act_ctrlrange'] = []
    for m in actuators:
        M['act_jntid'].append(M['jnt_name'].index(m.get('joint')))
        M['act_gear'].append(float(m.get('gear')))
        assert m.get('ctrllimited') == 'true'
        M['act_ctrlrange'].append(_floats(m.get('ctrlrange')))
    M['nq'], M['nv'], M['nu'] = nq[0], nv[0], len(actuators)
    M['nbody'], M['njnt'], M['ngeom'] = len(M['body_parent']), len(M['jnt_type']), len(M['geom_type'])
    M['timestep'] = timestep
    M['gravity'] = [0.0, 0.0, -9.81]
    M['scopes'] = scopes
    M['agent_names'] = list(agent_names)
    return M


def agent_slices(M):
    """Per-agent index bookkeeping of agents.py:45-83 (prefix match, contiguous slices)."""
    out = []
    for scope in M['scopes']:
        bodies = [i for i, n in enumerate(M['body_name']) if n.startswith(scope)]
        joints = [i for i, n in enumerate(M['jnt_name']) if n.startswith(scope)]
        q0 = M['jnt_qposadr'][joints[0]]
        q1 = M['jnt_qposadr'][joints[-1]] + (7 if M['jnt_type'][joints[-1]] == JNT_FREE else 1)
        dofs = [M['body_dofadr'][b] for b in bodies if M['body_dofadr'][b] >= 0]
        last = len(bodies) - 1
        while M['body_dofnum'][bodies[last]] == 0:
            last -= 1
        v0, v1 = dofs[0], dofs[-1] + M['body_dofnum'][bodies[last]]
        out.append(dict(bodies=bodies, q0=q0, q1=q1, v0=v0, v1=v1,
                        torso=[b for b in bodies if M['body_name'][b].endswith('/torso')][0]))
    return out
