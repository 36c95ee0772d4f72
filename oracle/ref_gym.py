"""CPU restatement of the reference's Python gym layer -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this.  It restates, around any engine object with the oracle ABI (oracle.ref_engine.RefEngine = the
compiled reference engine):
  gym_city/envs/tilemap.py:8-23, 65-199, 327-541   zoneFromInt, TileMap (shadow map, clear-before-build)
  gym_city/envs/corecontrol.py:36-357              MicropolisControl (window offset, tool tables, density maps)
  gym_city/envs/env.py:209-219, 253-292, 301-540   MicropolisEnv action decode, reset, observation, reward, step
The reference modules themselves need gym / GTK / the SWIG module and cannot be imported as a whole;
TileMap alone can (with gym stubbed) and tests/test_gym_oracle.py pins RefTileMap against it.
Road-network labelling (tilemap.py:201-330) is not observable through obs / reward / info and is omitted.
"""
import bisect

import numpy as np

ZONE_BPS = [('Land', None), ('Water', 1), ('Forest', 22), ('Rubble', 44), ('Flood', 48), ('Radioactive', 52),
            ('Trash', 53), ('Fire', 56), ('Bridge', 64), ('Road', 66), ('RoadWire', 77), ('Road', 79),
            ('Wire', 208), ('RailWire', 221), ('Trash', 223), ('Rail', 224), ('RoadRail', 237),
            ('RoadRail', 238), ('Residential', 240), ('Hospital', 405), ('Church', 410), ('Commercial', 423),
            ('Industrial', 610), ('Seaport', 693), ('Airport', 709), ('Radar', 711), ('Airport', 709),
            ('CoalPowerPlant', 745), ('FireDept', 761), ('PoliceDept', 770), ('Stadium', 779),
            ('NuclearPowerPlant', 811), ('Lightning', 827), ('Bridge', 828), ('Radar', 832), ('Park', 840),
            ('Net', 844), ('Industrial', 852), ('Rubble', 860), ('Industrial', 888), ('CoalPowerPlant', 916),
            ('Stadium', 932), ('Bridge', 948), ('NuclearPowerPlant', 952), ('Church', 956)]
_ZNAMES = [z[0] for z in ZONE_BPS]
_ZBREAKS = [z[1] for z in ZONE_BPS[1:]]


def zone_from_int(i):
    """tilemap.py:8-23 (bisect over a list that is not monotonic at 'Airport', 709)."""
    return _ZNAMES[bisect.bisect(_ZBREAKS, i)]


# Oracle pin: zoneFromInt can return 'Flood', 'Radioactive', 'Trash' and 'Lightning', which have no
# zoneSize / zoneCols entry, so the reference TileMap raises KeyError (and the env process dies) the
# first time a flooded or radioactive tile is re-read (tilemap.py:488-491, 455-456).  A batched env
# cannot die per env, so both sides classify such a tile as 'Rubble' and raise a sticky
# `tilemap_error` flag instead.  RefTileMap(strict=True) keeps the reference's KeyError.
UNSIZED = ('Flood', 'Radioactive', 'Trash', 'Lightning')


ENGINE_TOOLS = ['Residential', 'Commercial', 'Industrial', 'FireDept', 'PoliceDept', 'Query', 'Wire', 'Clear',
                'Rail', 'Road', 'Stadium', 'Park', 'Seaport', 'CoalPowerPlant', 'NuclearPowerPlant', 'Airport',
                'Net', 'Water', 'Land', 'Forest']                                   # corecontrol.py:61-80
AGENT_TOOLS = ['Residential', 'Commercial', 'Industrial', 'FireDept', 'PoliceDept', 'Clear', 'Wire', 'Rail',
               'Road', 'Stadium', 'Park', 'Seaport', 'CoalPowerPlant', 'NuclearPowerPlant', 'Airport', 'Net',
               'Water', 'Land', 'Forest', 'Nil']                                    # corecontrol.py:85-106


class RefTileMap:
    def __init__(self, micro, MAP_X, MAP_Y, strict=False):
        self.micro, self.MAP_X, self.MAP_Y = micro, MAP_X, MAP_Y
        self.strict, self.tilemap_error = strict, 0
        self.zoneSize = {'Residential': 3, 'Commercial': 3, 'Industrial': 3, 'Seaport': 4, 'Stadium': 4,
                         'PoliceDept': 3, 'FireDept': 3, 'Airport': 6, 'NuclearPowerPlant': 4,
                         'CoalPowerPlant': 4, 'Road': 1, 'Rail': 1, 'Park': 1, 'Wire': 1, 'Rubble': 1, 'Net': 1,
                         'Water': 1, 'Land': 1, 'Forest': 1, 'Church': 3, 'Hospital': 3, 'Fire': 1}
        link = {'Forest': ['Forest', 'Land'], 'Church': ['Church', 'Residential'],
                'Hospital': ['Hospital', 'Residential']}
        comp = {"RoadWire": ["Road", "Wire"], "RailWire": ["Rail", "Wire"], "Bridge": ["Road", "Water"],
                "RoadRail": ["Road", "Rail"], "WaterWire": ["Water", "Wire"], "Radar": ["Net", "Airport"],
                "RailBridge": ["Rail", "Water"]}
        compat = {}
        for k in comp:                      # tilemap.py:127-135 -- lists are shared on purpose (aliasing)
            l = comp[k]
            for c in l:
                if c in compat:
                    compat[c] += [z for z in l if z not in compat[c]]
                else:
                    compat[c] = l
        self.composite_zones, self.zone_compat = comp, compat
        self.zones = list(self.zoneSize.keys()) + list(comp.keys())
        for c in comp:
            self.zoneSize[c] = self.zoneSize[comp[c][0]]
        self.num_zones = len(self.zones)
        self.num_features = self.num_zones - len(comp)
        self.zoneInts = {z: i for i, z in enumerate(self.zones)}
        self.zoneCols = {}
        for z in self.zones:
            col = np.zeros(self.num_features + 1, dtype=np.uint8)
            feats = comp.get(z, link.get(z, [z]))
            for f in feats:
                col[self.zoneInts[f]] = 1
            col[-1] = self.zoneInts[z]
            self.zoneCols[z] = col
        self.zoneMap = np.zeros((self.num_features + 1, MAP_X, MAP_Y), dtype=np.uint8)
        self.static_builds = np.zeros((1, MAP_X, MAP_Y), dtype=int)
        self.centers = np.full((MAP_X, MAP_Y), None)
        self.n_struct_tiles = 0
        self.track_ages = False
        self.paint = False                 # tilemap.py:74-81 (MicropolisPaintControl passes paint=True)
        self.acted = np.zeros((MAP_X, MAP_Y))
        self.setEmpty()

    def init_age_array(self):                                       # tilemap.py:201-206
        self.track_ages = True
        self.age_order = np.full((self.MAP_X, self.MAP_Y), -1, dtype=int)

    def setEmpty(self):                                             # tilemap.py:327-343
        self.num_roads = 0
        self.num_plants = 0
        self.static_builds.fill(0)
        self.centers.fill(None)
        self.zoneMap.fill(0)
        self.zoneMap[-1, :, :] = self.zoneInts['Land']
        self.zoneMap[self.zoneInts['Land'], :, :] = 1

    def addZoneBot(self, x, y, tool, static_build=False):           # tilemap.py:345-377
        if tool == 'Nil':
            return True
        zone = tool
        old_zone = self.zones[self.zoneMap[-1][x][y]]
        if (zone in self.zone_compat and (old_zone in self.zone_compat[zone] or (
                old_zone in self.composite_zones and zone in self.composite_zones[old_zone]))) and zone != 'Water':
            if self.static_builds[0][x][y] == 1:
                static_build = True
        if (not static_build) and self.static_builds[0][x][y] == 1:
            return False
        if zone == 'Clear':
            zone = 'Land'
        result = self.clearForZone(x, y, zone, static_build=static_build)
        if result == 1:
            result = self.micro.doSimTool(x, y, tool)
            if result == 1:
                result = self.addZone(x, y, zone, static_build)
        return result == 1

    def clearForZone(self, x, y, zone, static_build=False):         # tilemap.py:382-389
        old_zone = self.zones[self.zoneMap[-1][x][y]]
        if zone in self.zone_compat and (old_zone in self.zone_compat[zone]):
            return 1
        return self.clearPatch(x, y, self.zoneSize[zone], static_build)

    def clearPatch(self, x, y, patch_size, static_build=False):     # tilemap.py:391-408
        if patch_size == 1 and not ((static_build is False) and self.static_builds[0][x][y] == 1):
            self.clearTile(x, y, static_build=static_build)
            return 1
        x0, x1 = max(x - 1, 0), min(x - 1 + patch_size, self.MAP_X)
        y0, y1 = max(y - 1, 0), min(y - 1 + patch_size, self.MAP_Y)
        for i in range(x0, x1):
            for j in range(y0, y1):
                if self.static_builds[0][i][j] == 1:
                    return 0
                if self.paint and self.acted[i][j] == 1:           # :402-404
                    return 0
                self.clearTile(i, j, static_build=static_build)
        return 1

    def clearTile(self, x, y, static_build=False):                  # tilemap.py:410-447
        old_zone = self.zones[self.zoneMap[-1][x][y]]
        cnt = self.centers[x][y]
        if cnt is None:
            cnt = (x, y)
        if self.static_builds[0][x][y] == 1 and static_build is False:
            return -1
        xc, yc = cnt[0], cnt[1]
        self.micro.doSimTool(xc, yc, 'Water')
        self.micro.doSimTool(xc, yc, 'Land')
        self.micro.doSimTool(xc, yc, 'Clear')
        size = self.zoneSize[old_zone]
        if size == 1:
            self.updateTile(xc, yc, static_build=False)
            return
        if size == 6:
            xc -= 1
            yc -= 1
        x0, y0 = max(0, xc - 1), max(0, yc - 1)
        x1, y1 = min(xc - 1 + size, self.MAP_X), min(yc - 1 + size, self.MAP_Y)
        for i in range(x0, x1):
            for j in range(y0, y1):
                self.updateTile(i, j, static_build=False)

    def _zone_at(self, x, y):
        zone = zone_from_int(self.micro.getTile(x, y))
        if zone in UNSIZED and not self.strict:
            self.tilemap_error = 1
            zone = 'Rubble'
        return zone

    def addZone(self, x, y, zone, static_build=False):              # tilemap.py:450-479
        zone = self._zone_at(x, y)
        zone_size = self.zoneSize[zone]
        center = (x + 1, y + 1) if zone_size == 6 else (x, y)
        if zone_size == 1:
            self.updateTile(x, y, zone, center, static_build)
            if self.paint:
                self.acted[x, y] = 1                                   # :465-466
            if self.track_ages:
                self.age_order[x][y] = self.micro.env.num_step         # :467-468
            return
        x0, y0 = max(0, x - 1), max(0, y - 1)
        x1, y1 = min(x - 1 + zone_size, self.MAP_X), min(y - 1 + zone_size, self.MAP_Y)
        for i in range(x0, x1):
            for j in range(y0, y1):
                self.updateTile(i, j, zone, center, static_build)
                if self.paint:
                    self.acted[i, j] = 1                               # :476-477
                if self.track_ages:
                    self.age_order[i][j] = self.micro.env.num_step     # :478-479

    def _flags(self, x, y):
        zi = self.zoneInts
        zm = self.zoneMap
        plant = zm[zi['NuclearPowerPlant']][x][y] == 1 or zm[zi['CoalPowerPlant']][x][y] == 1
        road = zm[zi['Road']][x][y] == 1
        natural = (zm[zi['Land']][x][y] == 1 or zm[zi['Water']][x][y] == 1 or zm[zi['Rubble']][x][y] == 1
                   or zm[zi['Forest']][x][y] == 1)
        return plant, road, natural

    def updateTile(self, x, y, zone=None, center=None, static_build=None):   # tilemap.py:467-541
        was_plant, was_road, was_natural = self._flags(x, y)
        if zone is None:
            zone = self._zone_at(x, y)
            center = self.centers[x][y] if self.zoneSize[zone] != 1 else (x, y)
            if static_build is False:
                self.static_builds[0][x][y] = 0
        else:
            if static_build and zone not in ['Land', 'Rubble', 'Water']:
                self.static_builds[0][x][y] = 1
            else:
                self.static_builds[0][x][y] = 0
        self.zoneMap[:, x, y] = self.zoneCols[zone]
        self.centers[x][y] = center
        is_plant, is_road, is_natural = self._flags(x, y)
        if was_natural and not is_natural:
            self.n_struct_tiles += 1
        if not was_natural and is_natural:
            self.n_struct_tiles -= 1
        if self.track_ages and is_natural:
            self.age_order[x][y] = -1                                  # :521-523
        if was_road and not is_road:
            self.num_roads -= 1
        elif not was_road and is_road:
            self.num_roads += 1
        if was_plant and not is_plant:
            self.num_plants -= 1
        elif not was_plant and is_plant:
            self.num_plants += 1

    def getMapState(self):
        return self.zoneMap[:-1, :, :]


class RefControl:
    """corecontrol.py:36-357 around an oracle-ABI engine."""

    def __init__(self, engine, MAP_W, MAP_H, power_puzzle=False, xs=16, ys=8, tilemap_cls=RefTileMap):
        self.engine = engine
        self.MAP_X, self.MAP_Y, self.MAP_XS, self.MAP_YS = MAP_W, MAP_H, xs, ys
        self.engineTools = ENGINE_TOOLS
        self.tools = ['Wire'] if power_puzzle else AGENT_TOOLS
        self.num_tools = len(self.tools)
        self.map = tilemap_cls(self, MAP_W, MAP_H)
        self.init_funds = 2000000
        self.total_traffic = 0

    def simTick(self):                                              # :148-152
        self.engine.controlSimTick()

    def doBotTool(self, x, y, tool, static_build=False):            # :270-272 (PADDING = 0)
        return self.map.addZoneBot(x, y, tool, static_build=static_build)

    def doSimTool(self, x, y, tool):                                # :296-301, 307-311
        return self.engine.toolDown(self.engineTools.index(tool), x + self.MAP_XS, y + self.MAP_YS)

    def getTile(self, x, y):                                        # :302-305
        return self.engine.getTile(x + self.MAP_XS, y + self.MAP_YS) & 1023

    def updateMap(self):                                            # :193-203
        for i in range(self.MAP_X):
            for j in range(self.MAP_Y):
                zone = self.map._zone_at(i, j) if hasattr(self.map, '_zone_at') else zone_from_int(self.getTile(i, j))
                self.map.updateTile(i, j, zone, (i, j))

    def clearMap(self):                                             # :188-192
        self.map.setEmpty()
        self.engine.clearMap()
        self.updateMap()

    def getDensityMaps(self):                                       # :215-239
        self.total_traffic = 0
        d = np.zeros((3, self.MAP_X, self.MAP_Y))
        traffic = self.engine.get("MPF_TRAFFIC_DENSITY").reshape(60, 50)
        pop = self.engine.get("MPF_POP_DENSITY").reshape(60, 50)
        power = self.engine.get("MPF_POWER_GRID").reshape(120, 100)

        def half(m, x, y):
            return int(m[x, y]) if (0 <= x < 60 and 0 <= y < 50) else 0
        for i in range(self.MAP_X // 2):
            for j in range(self.MAP_Y // 2):
                im, jm = i + self.MAP_YS // 2, j + self.MAP_XS // 2
                t = half(traffic, jm, im)
                self.total_traffic += t
                d[2][2 * i:2 * i + 2, 2 * j:2 * j + 2] = t
                d[1][2 * i:2 * i + 2, 2 * j:2 * j + 2] = half(pop, jm, im)
        for i in range(self.MAP_X):
            for j in range(self.MAP_Y):
                x, y = j + self.MAP_XS, i + self.MAP_YS
                d[0][i][j] = int(power[x, y]) if (0 <= x < 120 and 0 <= y < 100) else 0
        return d

    def takeAction(self, a, static_build=False):                    # :332-338
        return self.map.addZoneBot(int(a[1]), int(a[2]), self.tools[a[0]], static_build=static_build)


class RefEnv:
    """env.py MicropolisEnv step / reset / observation / reward around RefControl."""

    def __init__(self, engine, size, max_step=None, empty_start=True, power_puzzle=False, xs=16, ys=8,
                 tilemap_cls=RefTileMap):
        self.MAP_X, self.MAP_Y = (size, size) if isinstance(size, int) else size
        self.max_step = self.MAP_X * self.MAP_Y if max_step is None else max_step
        self.empty_start = empty_start
        self.micro = RefControl(engine, self.MAP_X, self.MAP_Y, power_puzzle, xs, ys, tilemap_cls)
        self.micro.env = self                                       # TileMap reads micro.env.num_step for ages
        self.city_trgs = {'res_pop': 500, 'com_pop': 50, 'ind_pop': 50, 'traffic': 2000, 'num_plants': 14,
                          'mayor_rating': 100}
        self.weights = {'res_pop': 1, 'com_pop': 1, 'ind_pop': 1, 'traffic': 1, 'num_plants': 0,
                        'mayor_rating': 0}
        self.num_tools = self.micro.num_tools
        self.num_step = 0
        self.poet = False
        self.param_ranges = [750, 100, 100, 2000, 100, 100]         # env.py:39-48 param_bounds, ub - lb
        self.city_metrics = self.get_city_metrics()
        self.last_city_metrics = self.city_metrics

    def decode(self, a):                                            # env.py:209-219
        n = self.MAP_X * self.MAP_Y
        return [a // n, (a % n) // self.MAP_Y, a % self.MAP_Y]

    def get_city_metrics(self):                                     # env.py:426-440
        s = self.micro.engine.scalars()
        return {'res_pop': s.resPop, 'com_pop': s.comPop, 'ind_pop': s.indPop,
                'traffic': self.micro.total_traffic, 'num_plants': self.micro.map.num_plants,
                'mayor_rating': s.cityYes}

    def getReward(self):                                            # env.py:345-387
        reward = 0
        for metric, trg in self.city_trgs.items():
            last_val = self.last_city_metrics[metric]
            trg_change = trg - last_val
            change = self.city_metrics[metric] - last_val
            if np.sign(change) != np.sign(trg_change):
                r = -abs(change)
            elif abs(change) < abs(trg_change):
                r = abs(change)
            else:
                r = abs(trg_change) - abs(trg_change - change)
            reward += r * self.weights[metric]
        return reward

    def getState(self):                                             # env.py:301-343
        s = self.micro.engine.scalars()
        f = np.float32
        scalars = [s.resPop, s.comPop, s.indPop, float(f(s.resValve) / f(2000)), float(f(s.comValve) / f(1500)),
                   float(f(s.indValve) / f(1500))]
        if getattr(self, 'poet', False):                            # env.py:305-312
            for j in range(3):
                scalars[j] = scalars[j] / self.param_ranges[j]
            trg_metrics = [v for k, v in self.city_trgs.items()]
            for i in range(len(trg_metrics)):
                trg_metrics[i] = trg_metrics[i] / self.param_ranges[i]
            scalars += trg_metrics
        state = self.micro.map.getMapState()
        density = self.micro.getDensityMaps()
        layers = np.zeros((len(scalars), self.MAP_X, self.MAP_Y))
        for i, v in enumerate(scalars):
            layers[i].fill(v)
        return np.concatenate((state, density, layers, self.micro.map.static_builds), 0)

    def reset_begin(self, terrain_seed=None, sim_seed=None):        # env.py:264-274
        self.micro.clearMap()
        if not self.empty_start:
            self.micro.engine.generate(terrain_seed)
            self.micro.engine.seed(sim_seed)
            self.micro.updateMap()
        self.num_step = 0

    def random_static_start(self, k, actions):                      # env.py:228-241 with the draws passed in
        self.micro.engine.setFunds(self.micro.init_funds)
        for j in range(int(k)):
            self.step(int(actions[j]), static_build=True)

    def reset_finish(self):                                         # env.py:279-292
        self.micro.simTick()
        self.city_metrics = self.get_city_metrics()
        self.last_city_metrics = self.city_metrics
        self.micro.engine.setFunds(self.micro.init_funds)
        self.curr_reward = self.getReward()
        self.state = self.getState()
        return self.state

    def step(self, a, static_build=False):                          # env.py:448-540
        if a >= 0:
            self.micro.takeAction(self.decode(a), static_build)
        self.micro.engine.setFunds(self.micro.init_funds)
        self.micro.simTick()
        self.state = self.getState()
        self.last_city_metrics = self.city_metrics
        self.city_metrics = self.get_city_metrics()
        reward = self.getReward()
        funds = self.micro.engine.scalars().totalFunds
        terminal = (funds < 0 or self.num_step >= self.max_step)
        self.num_step += 1
        return self.state, reward, terminal, {}


class RefExtinguisher:
    """gym_city/wrappers.py:10-141 Extinguisher around a RefEnv.  The reference draws from numpy's global
    RNG (localWipe's centre, ranDemolish's np.random.choice); here the caller passes the draws
    (`rnd`: n_dels + 2 non-negative ints per event) so that both sides of a parity test use the same
    ones: ranDemolish takes aged cell number rnd[i] % count, localWipe the centre (rnd[0] % MAP_X,
    rnd[1] % MAP_Y)."""
    TYPES = ('age', 'spatial', 'random', 'Monster')

    def __init__(self, env, extinction_type=None, extinction_prob=0.1, xt_dels=25):
        self.env = env
        self.set_extinction_type(extinction_type, extinction_prob, xt_dels)

    def set_extinction_type(self, extinction_type, extinction_prob, extinction_dels):   # :24-33
        self.extinction_type, self.extinction_prob, self.n_dels = extinction_type, extinction_prob, extinction_dels
        self.extinction_interval = -1 if extinction_prob == 0 else 1 / extinction_prob
        self.env.micro.map.init_age_array()

    def step(self, a, rnd=None, static_build=False):                # :35-40
        out = self.env.step(a, static_build)
        dels = None
        if self.env.num_step % self.extinction_interval == 0:
            dels = self.extinguish(self.extinction_type, rnd)
        return out, dels

    def extinguish(self, extinction_type='age', rnd=None):          # :42-55
        env, m = self.env, self.env.micro
        if extinction_type == 'Monster':
            m.engine.makeMonster()
            return 0
        ages = lambda: m.map.age_order
        curr_dels = 0
        if extinction_type == 'age':                                # elderCleanse :118-141
            for i in range(self.n_dels):
                a = ages().flatten()
                youngest = np.max(a)
                if len(np.where(a > -1)[0]) == 0:
                    break
                a = np.copy(a)
                a += (a < 0) * 2 * youngest
                x, y = np.unravel_index(np.argmin(a), ages().shape)
                m.doBotTool(int(x), int(y), 'Clear', static_build=True)
                curr_dels += 1
            m.engine.setFunds(m.init_funds)
        elif extinction_type == 'random':                           # ranDemolish :88-104
            for i in range(self.n_dels):
                a = ages().flatten()
                age_is = np.where(a > -1)[0]
                if len(age_is) == 0:
                    break
                age_i = age_is[int(rnd[i]) % len(age_is)]
                x, y = np.unravel_index(age_i, ages().shape)
                m.doBotTool(int(x), int(y), 'Clear', static_build=True)
                curr_dels += 1
        elif extinction_type == 'spatial':                          # localWipe :57-67, clear_border :69-86
            x, y = int(rnd[0]) % env.MAP_X, int(rnd[1]) % env.MAP_Y
            for r in range(0, max(x, abs(env.MAP_X - x), y, abs(env.MAP_Y - y))):
                done = False
                for x_i in range(x - r, x + r):
                    if x_i < 0 or x_i >= env.MAP_X:
                        continue
                    for y_i in range(y - r, y + r):
                        if y_i < 0 or y_i >= env.MAP_X or y_i >= env.MAP_Y:
                            continue
                        if ages()[x_i, y_i] > 0:
                            m.doBotTool(x_i, y_i, 'Clear', static_build=True)
                            curr_dels += 1
                            if curr_dels == self.n_dels:
                                done = True
                                break
                    if done:
                        break
                if curr_dels >= self.n_dels:
                    break
        return curr_dels


class RefPaintEnv(RefEnv):
    """gym_city/envs/paintenv.py:20-53 + paintcontrol.py:36-70 around the oracle control: step(a) with a (W, H)
    array of tool indices, one doBotTool per cell (the last column is never reached), cells an earlier build
    of the same step covers are skipped.  The TileMap's paint rule is switched on for the action only (setup
    builds of the reference happen with it on, which differs for static cells only, and those refuse
    non-static builds anyway)."""

    def step(self, a, static_build=False):
        m = self.micro
        a = np.asarray(a)
        m.map.paint = True
        i = 0
        while i in range(self.MAP_X):
            j = 0
            while j in range(self.MAP_Y):
                if j >= self.MAP_Y - 1 or i >= self.MAP_X:
                    break
                tool = m.tools[int(a[i][j])]
                if m.map.acted[i, j] == 0:
                    m.doBotTool(i, j, tool, static_build=False)
                j += 1
            i += 1
        m.map.acted.fill(0)
        m.map.paint = False
        return RefEnv.step(self, -1, static_build)          # postact (env.py:468-540); the reward is recomputed there
