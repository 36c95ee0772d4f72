// mp_sprites.h -- moving objects (train, helicopter, airplane, ship, monster, tornado, explosion)
// and disasters.  All of it is the RNG-ordered serial lane: [LEAD] only.
//
// The reference keeps sprites in a singly linked list with PREPEND on create and a free pool
// (sprite.cpp:93-118, 303-329), so iteration order = reverse creation order, and makeSprite()
// re-initialises the per-type "global" sprite in place (keeping its list position).  Here the
// list lives in a fixed table of MAX_SPRITES slots linked by indices, which reproduces that
// order exactly; a full table drops the new sprite and raises Scalars::spriteOverflow.
#pragma once
#include "mp_sim.h"

namespace mp {
#define S (*e.s)

MPD int get_sprite(Env &e, int type)                         // sprite.cpp:336-344
{
    MP_COV(get_sprite);
    int i = S.globalSprites[type];
    if (i < 0 || e.spr()[i].frame == 0) return -1;
    return i;
}

// [LEAD] sprite.cpp:125-296 initSprite
MPN void init_sprite(Env &e, int si, int x, int y)
{
    MP_COV(init_sprite);
    Sprite &sp = e.spr()[si];
    sp.x = (int16_t)x; sp.y = (int16_t)y; sp.frame = 0;
    sp.origX = 0; sp.origY = 0; sp.destX = 0; sp.destY = 0; sp.count = 0; sp.soundCount = 0;
    sp.dir = 0; sp.newDir = 0; sp.step = 0; sp.flag = 0; sp.control = -1; sp.turn = 0;
    sp.accel = 0; sp.speed = 100;
    if (S.globalSprites[sp.type] < 0) S.globalSprites[sp.type] = (int16_t)si;
    switch (sp.type) {
    case SPR_TRAIN:
        sp.width = 32; sp.height = 32; sp.xOffset = 32; sp.yOffset = -16; sp.xHot = 40; sp.yHot = -8;
        sp.frame = 1; sp.dir = 4;
        break;
    case SPR_SHIP:
        sp.width = 48; sp.height = 48; sp.xOffset = 32; sp.yOffset = -16; sp.xHot = 48; sp.yHot = 0;
        if (x < (4 << 4)) sp.frame = 3;
        else if (x >= ((WORLD_W - 4) << 4)) sp.frame = 7;
        else if (y < (4 << 4)) sp.frame = 5;
        else if (y >= ((WORLD_H - 4) << 4)) sp.frame = 1;
        else sp.frame = 3;
        sp.newDir = sp.frame; sp.dir = 10; sp.count = 1;
        break;
    case SPR_MONSTER:
        sp.width = 48; sp.height = 48; sp.xOffset = 24; sp.yOffset = 0; sp.xHot = 40; sp.yHot = 16;
        if (x > ((WORLD_W << 4) / 2)) sp.frame = (y > ((WORLD_H << 4) / 2)) ? 10 : 7;
        else if (y > ((WORLD_H << 4) / 2)) sp.frame = 1;
        else sp.frame = 4;
        sp.count = 1000;
        sp.destX = (int16_t)(S.pollutionMaxX << 4); sp.destY = (int16_t)(S.pollutionMaxY << 4);
        sp.origX = sp.x; sp.origY = sp.y;
        break;
    case SPR_HELICOPTER:
        sp.width = 32; sp.height = 32; sp.xOffset = 32; sp.yOffset = -16; sp.xHot = 40; sp.yHot = -8;
        sp.frame = 5; sp.count = 1500;
        sp.destX = rnd(e, (WORLD_W << 4) - 1);
        sp.destY = rnd(e, (WORLD_H << 4) - 1);
        sp.origX = (int16_t)(x - 30); sp.origY = (int16_t)y;
        break;
    case SPR_AIRPLANE:
        sp.width = 48; sp.height = 48; sp.xOffset = 24; sp.yOffset = 0; sp.xHot = 48; sp.yHot = 16;
        if (x > ((WORLD_W - 20) << 4)) {
            sp.x = (int16_t)(sp.x - (100 + 48));
            sp.destX = (int16_t)(sp.x - 200);
            sp.frame = 7;
        } else {
            sp.destX = (int16_t)(sp.x + 200);
            sp.frame = 11;
        }
        sp.destY = sp.y;
        break;
    case SPR_TORNADO:
        sp.width = 48; sp.height = 48; sp.xOffset = 24; sp.yOffset = 0; sp.xHot = 40; sp.yHot = 36;
        sp.frame = 1; sp.count = 200;
        break;
    case SPR_EXPLOSION:
        sp.width = 48; sp.height = 48; sp.xOffset = 24; sp.yOffset = 0; sp.xHot = 40; sp.yHot = 16;
        sp.frame = 1;
        break;
    case SPR_BUS:
        sp.width = 32; sp.height = 32; sp.xOffset = 30; sp.yOffset = -18; sp.xHot = 40; sp.yHot = -8;
        sp.frame = 1; sp.dir = 1;
        break;
    default: break;
    }
}

// [LEAD] sprite.cpp:93-118 newSprite
MPN int new_sprite(Env &e, int type, int x, int y)
{
    MP_COV(new_sprite);
    int si = -1;
    MP_NOUNROLL for (int i = 0; i < MAX_SPRITES; i++)
        if (!e.spr()[i].used) { si = i; break; }
    if (si < 0) { S.spriteOverflow = 1; return -1; }
    e.spr()[si].used = 1;
    e.spr()[si].type = (int16_t)type;
    init_sprite(e, si, x, y);
    e.spr()[si].next = S.spriteHead;
    S.spriteHead = (int16_t)si;
    return si;
}

// [LEAD] sprite.cpp:303-329 destroySprite
MPD void destroy_sprite(Env &e, int si)
{
    MP_COV(destroy_sprite);
    if (S.globalSprites[e.spr()[si].type] == si) S.globalSprites[e.spr()[si].type] = -1;
    if (S.spriteHead == si) S.spriteHead = e.spr()[si].next;
    else
        MP_NOUNROLL for (int p = S.spriteHead; p >= 0; p = e.spr()[p].next)
            if (e.spr()[p].next == si) { e.spr()[p].next = e.spr()[si].next; break; }
    e.spr()[si].used = 0;
}

// [LEAD] sprite.cpp:352-365 makeSprite
MPD int make_sprite(Env &e, int type, int x, int y)
{
    MP_COV(make_sprite);
    int si = S.globalSprites[type];
    if (si < 0) return new_sprite(e, type, x, y);
    init_sprite(e, si, x, y);
    return si;
}

MPN void make_explosion_at(Env &e, int x, int y) { new_sprite(e, SPR_EXPLOSION, x - 40, y - 16); }   // sprite.cpp:2029
MPN void make_explosion(Env &e, int x, int y)                  // sprite.cpp:2017-2022
{
    MP_COV(make_explosion);
    if (inb(x, y)) make_explosion_at(e, (x << 4) + 8, (y << 4) + 8);
}

MPD int boat_distance(Env &e, int x, int y)                    // simulate.cpp:1253-1273
{
    MP_COV(boat_distance);
    int dist = 99999, mx = x * 16 + 8, my = y * 16 + 8;
    MP_NOUNROLL for (int i = S.spriteHead; i >= 0; i = e.spr()[i].next) {
        const Sprite &sp = e.spr()[i];
        if (sp.type == SPR_SHIP && sp.frame != 0) {
            int d = iabs(sp.x + sp.xHot - mx) + iabs(sp.y + sp.yHot - my);
            dist = tmin(dist, d);
        }
    }
    return dist;
}

MPD int16_t get_char(const Env &e, int x, int y)               // sprite.cpp:376-386
{
    MP_COV(get_char);
    x >>= 4; y >>= 4;
    if (!inb(x, y)) return -1;
    return (int16_t)(e.map()[tix(x, y)] & LOMASK);
}

MPD int16_t turn_to(int p, int d)                              // sprite.cpp:395-423
{
    MP_COV(turn_to);
    if (p == d) return (int16_t)p;
    if (p < d) { if (d - p < 4) p++; else p--; }
    else { if (p - d < 4) p--; else p++; }
    if (p > 8) p = 1;
    if (p < 1) p = 8;
    return (int16_t)p;
}

MPD bool try_other(int tpoo, int told, int tnew)               // sprite.cpp:426-443
{
    MP_COV(try_other);
    int16_t z = (int16_t)(told + 4);
    if (z > 8) z -= 8;
    if (tnew != z) return false;
    return tpoo == T_POWERBASE || tpoo == T_POWERBASE + 1 || tpoo == T_RAILBASE || tpoo == T_RAILBASE + 1;
}

MPD bool sprite_not_in_bounds(const Sprite &sp)                // sprite.cpp:451-457
{
    MP_COV(sprite_not_in_bounds);
    int x = sp.x + sp.xHot, y = sp.y + sp.yHot;
    return x < 0 || y < 0 || x >= (WORLD_W << 4) || y >= (WORLD_H << 4);
}

// sprite.cpp:471-511 getDir (sets absDist; keeps the `dispY * 2 < dispY` slip)
MPD int16_t get_dir(Env &e, int ox, int oy, int dx_, int dy_)
{
    MP_COV(get_dir);
    MP_TAB int16_t gd[13] = { 0, 3, 2, 1, 3, 4, 5, 7, 6, 5, 7, 8, 1 };
    int dispX = dx_ - ox, dispY = dy_ - oy, z;
    if (dispX < 0) z = dispY < 0 ? 11 : 8;
    else z = dispY < 0 ? 2 : 5;
    dispX = iabs(dispX); dispY = iabs(dispY);
    S.absDist = dispX + dispY;
    if (dispX * 2 < dispY) z++;
    else if (dispY * 2 < dispY) z--;
    if (z < 0 || z > 12) z = 0;
    return gd[z];
}

MPD bool sprite_collision(const Sprite &a, const Sprite &b)    // sprite.cpp:536-541
{
    MP_COV(sprite_collision);
    return a.frame != 0 && b.frame != 0
        && (iabs((a.x + a.xHot) - (b.x + b.xHot)) + iabs((a.y + a.yHot) - (b.y + b.yHot))) < 30;
}

MPD void explode_sprite(Env &e, int si)                        // sprite.cpp:1649-1688
{
    MP_COV(explode_sprite);
    Sprite &sp = e.spr()[si];
    sp.frame = 0;
    make_explosion_at(e, sp.x + sp.xHot, sp.y + sp.yHot);
}

// [LEAD] sprite.cpp:1752-1786 startFireInZone
MPD void start_fire_in_zone(Env &e, int x, int y, int ch)
{
    MP_COV(start_fire_in_zone);
    int v = tclamp(s8get(e.rog(), x, y) - 20, -200, 200);
    s8set(e.rog(), x, y, v);
    ch &= LOMASK;
    int xymax = ch < T_PORTBASE ? 2 : (ch == T_AIRPORT ? 5 : 4);
    MP_NOUNROLL for (int dx = -1; dx < xymax; dx++)
        MP_NOUNROLL for (int dy = -1; dy < xymax; dy++) {
            int xx = x + dx, yy = y + dy;
            if (inb(xx, yy) && (e.map()[tix(xx, yy)] & LOMASK) >= T_ROADBASE) e.map()[tix(xx, yy)] |= BULLBIT;
        }
}

// [LEAD] sprite.cpp:1709-1746 destroyMapTile
MPS void destroy_map_tile(Env &e, int ox, int oy)
{
    MP_COV(destroy_map_tile);
    int x = (int16_t)(ox >> 4), y = (int16_t)(oy >> 4);
    if (!inb(x, y)) return;
    int16_t z = (int16_t)e.map()[tix(x, y)];
    int16_t t = z & LOMASK;
    if (t >= T_TREEBASE) {
        if (!(z & BURNBIT)) {
            if (t >= T_ROADBASE && t <= T_LASTROAD) mset(e, tix(x, y), T_RIVER);
            return;
        }
        if (z & ZONEBIT) {
            start_fire_in_zone(e, x, y, z);
            if (t > T_RZB) make_explosion_at(e, ox, oy);
        }
        bool wet = t == T_HPOWER || t == T_VPOWER || t == T_HRAIL || t == T_VRAIL || t == T_BRWH || t == T_BRWV;
        if (wet) mset(e, tix(x, y), T_RIVER);
        else mset(e, tix(x, y), (uint16_t)((S.doAnimation ? (int)T_TINYEXP : (T_LASTTINYEXP - 3)) | BULLBIT | ANIMBIT));
    }
}

// [LEAD] sprite.cpp:1800-1822 startFire
MPD void start_fire(Env &e, int x, int y)
{
    MP_COV(start_fire);
    x >>= 4; y >>= 4;
    if (!inb(x, y)) return;
    int z = e.map()[tix(x, y)], t = z & LOMASK;
    if (!(z & BURNBIT) && t != T_DIRT) return;
    if (z & ZONEBIT) return;
    mset(e, tix(x, y), (uint16_t)random_fire(e));
}

// ------------------------------------------------------------------ per-type movement
// [LEAD] sprite.cpp:608-673 doTrainSprite
MPN void do_train(Env &e, int si)
{
    MP_COV(do_train);
    Sprite &sp = e.spr()[si];
    MP_TAB int16_t Cx[4] = { 0, 16, 0, -16 }, Cy[4] = { -16, 0, 16, 0 };
    MP_TAB int16_t Dx[5] = { 0, 4, 0, -4, 0 }, Dy[5] = { -4, 0, 4, 0, 0 };
    MP_TAB int16_t pic2[5] = { 1, 2, 1, 2, 5 };
    if (sp.frame == 3 || sp.frame == 4) sp.frame = pic2[sp.dir];
    sp.x = (int16_t)(sp.x + Dx[sp.dir]);
    sp.y = (int16_t)(sp.y + Dy[sp.dir]);
    if ((S.spriteCycle & 3) == 0) {
        int16_t dir = (int16_t)(rnd16(e) & 3);
        MP_NOUNROLL for (int16_t z = dir; z < dir + 4; z++) {
            int16_t dir2 = z & 3;
            if (sp.dir != 4 && dir2 == ((sp.dir + 2) & 3)) continue;
            int16_t c = get_char(e, sp.x + Cx[dir2] + 48, sp.y + Cy[dir2]);
            if ((c >= T_RAILBASE && c <= T_LASTRAIL) || c == T_RAILVPOWERH || c == T_RAILHPOWERV) {
                if (sp.dir != dir2 && sp.dir != 4) sp.frame = (sp.dir + dir2 == 3) ? 3 : 4;
                else sp.frame = pic2[dir2];
                if (c == T_HRAIL || c == T_VRAIL) sp.frame = 5;
                sp.dir = dir2;
                return;
            }
        }
        if (sp.dir == 4) { sp.frame = 0; return; }
        sp.dir = 4;
    }
}

// [LEAD] sprite.cpp:680-770 doCopterSprite
MPS void do_copter(Env &e, int si)
{
    MP_COV(do_copter);
    Sprite &sp = e.spr()[si];
    MP_TAB int16_t CDx[9] = { 0, 0, 3, 5, 3, 0, -3, -5, -3 }, CDy[9] = { 0, -5, -3, 0, 3, 5, 3, 0, -3 };
    if (sp.soundCount > 0) sp.soundCount--;
    if (sp.control < 0) {
        if (sp.count > 0) sp.count--;
        if (sp.count == 0) {
            int s2 = get_sprite(e, SPR_MONSTER);
            if (s2 < 0) s2 = get_sprite(e, SPR_TORNADO);
            if (s2 >= 0) { sp.destX = e.spr()[s2].x; sp.destY = e.spr()[s2].y; }
            else { sp.destX = sp.origX; sp.destY = sp.origY; }
            get_dir(e, sp.x, sp.y, sp.origX, sp.origY);
            if (S.absDist < 30) { sp.frame = 0; return; }
        }
    } else {
        get_dir(e, sp.x, sp.y, sp.destX, sp.destY);
        if (S.absDist < 16) { sp.destX = sp.origX; sp.destY = sp.origY; sp.control = -1; }
    }
    if (sp.soundCount == 0) {
        int16_t x = (int16_t)((sp.x + 48) / 16), y = (int16_t)(sp.y / 16);
        if (x >= 0 && x < WORLD_W && y >= 0 && y < WORLD_H)
            if (b2get(e.traffic(), x, y) > 170 && (rnd16(e) & 7) == 0) sp.soundCount = 200;
    }
    int16_t z = sp.frame;
    if ((S.spriteCycle & 3) == 0) {
        int16_t d = get_dir(e, sp.x, sp.y, sp.destX, sp.destY);
        z = turn_to(z, d);
        sp.frame = z;
    }
    sp.x = (int16_t)(sp.x + CDx[z]);
    sp.y = (int16_t)(sp.y + CDy[z]);
}

// [LEAD] sprite.cpp:777-850 doAirplaneSprite
MPS void do_airplane(Env &e, int si)
{
    MP_COV(do_airplane);
    MP_TAB int16_t CDx[12] = { 0, 0, 6, 8, 6, 0, -6, -8, -6, 8, 8, 8 };
    MP_TAB int16_t CDy[12] = { 0, -8, -6, 0, 6, 8, 6, 0, -6, 0, 0, 0 };
    int16_t z = e.spr()[si].frame;
    if ((S.spriteCycle % 5) == 0) {
        if (z > 8) {
            z--;
            if (z < 9) z = 3;
            e.spr()[si].frame = z;
        } else {
            int16_t d = get_dir(e, e.spr()[si].x, e.spr()[si].y, e.spr()[si].destX, e.spr()[si].destY);
            z = turn_to(z, d);
            e.spr()[si].frame = z;
        }
    }
    if (S.absDist < 50) {
        e.spr()[si].destX = (int16_t)(rnd(e, (WORLD_W * 16) + 100) - 50);
        e.spr()[si].destY = (int16_t)(rnd(e, (WORLD_H * 16) + 100) - 50);
    }
    if (S.enableDisasters) {
        bool explode = false;
        MP_NOUNROLL for (int s = S.spriteHead; s >= 0; s = e.spr()[s].next) {
            if (e.spr()[s].frame == 0 || s == si) continue;
            if ((e.spr()[s].type == SPR_HELICOPTER || e.spr()[s].type == SPR_AIRPLANE)
                && sprite_collision(e.spr()[si], e.spr()[s])) {
                explode_sprite(e, s);
                explode = true;
            }
        }
        if (explode) explode_sprite(e, si);
    }
    e.spr()[si].x = (int16_t)(e.spr()[si].x + CDx[z]);
    e.spr()[si].y = (int16_t)(e.spr()[si].y + CDy[z]);
    if (sprite_not_in_bounds(e.spr()[si])) e.spr()[si].frame = 0;
}

// [LEAD] sprite.cpp:857-984 doShipSprite
MPS void do_ship(Env &e, int si)
{
    MP_COV(do_ship);
    Sprite &sp = e.spr()[si];
    MP_TAB int16_t BDx[9] = { 0, 0, 1, 1, 1, 0, -1, -1, -1 }, BDy[9] = { 0, -1, -1, 0, 1, 1, 1, 0, -1 };
    MP_TAB int16_t BPx[9] = { 0, 0, 2, 2, 2, 0, -2, -2, -2 }, BPy[9] = { 0, -2, -2, 0, 2, 2, 2, 0, -2 };
    MP_TAB int16_t clr[8] = { T_RIVER, T_CHANNEL, T_POWERBASE, T_POWERBASE + 1, T_RAILBASE, T_RAILBASE + 1,
        T_BRWH, T_BRWV };
    int16_t x, y, z, t = T_RIVER;
    if (sp.soundCount > 0) sp.soundCount--;
    if (!sp.soundCount) {
        if ((rnd16(e) & 3) == 1) {
            // scenario is always SC_NONE here, so the San Francisco fog-horn draw never happens
        }
        sp.soundCount = 200;
    }
    if (sp.count > 0) sp.count--;
    if (sp.count == 0) {
        sp.count = 9;
        if (sp.frame != sp.newDir) {
            sp.frame = turn_to(sp.frame, sp.newDir);
            return;
        }
        int16_t tem = (int16_t)(rnd16(e) & 7), pem;
        MP_NOUNROLL for (pem = tem; pem < (tem + 8); pem++) {
            z = (int16_t)((pem & 7) + 1);
            if (z == sp.dir) continue;
            x = (int16_t)(((sp.x + (48 - 1)) >> 4) + BDx[z]);
            y = (int16_t)((sp.y >> 4) + BDy[z]);
            if (inb(x, y)) {
                t = (int16_t)(e.map()[tix(x, y)] & LOMASK);
                if (t == T_CHANNEL || t == T_BRWH || t == T_BRWV || try_other(t, sp.dir, z)) {
                    sp.newDir = z;
                    sp.frame = turn_to(sp.frame, sp.newDir);
                    sp.dir = (int16_t)(z + 4);
                    if (sp.dir > 8) sp.dir -= 8;
                    break;
                }
            }
        }
        if (pem == (tem + 8)) {
            sp.dir = 10;
            sp.newDir = (int16_t)((rnd16(e) & 7) + 1);
        }
    } else {
        z = sp.frame;
        if (z == sp.newDir) {
            sp.x = (int16_t)(sp.x + BPx[z]);
            sp.y = (int16_t)(sp.y + BPy[z]);
        }
    }
    if (sprite_not_in_bounds(sp)) { sp.frame = 0; return; }
    MP_NOUNROLL for (z = 0; z < 8; z++) {
        if (t == clr[z]) break;
        if (z == 7) {
            explode_sprite(e, si);
            destroy_map_tile(e, e.spr()[si].x + 48, e.spr()[si].y);
        }
    }
}

MPD void hit_vehicles(Env &e, int si)                          // shared by monster and tornado
{
    MP_COV(hit_vehicles);
    MP_NOUNROLL for (int s = S.spriteHead; s >= 0; s = e.spr()[s].next) {
        int ty = e.spr()[s].type;
        if (e.spr()[s].frame != 0
            && (ty == SPR_AIRPLANE || ty == SPR_HELICOPTER || ty == SPR_SHIP || ty == SPR_TRAIN)
            && sprite_collision(e.spr()[si], e.spr()[s]))
            explode_sprite(e, s);
    }
}

// [LEAD] sprite.cpp:997-1181 doMonsterSprite
MPN void do_monster(Env &e, int si)
{
    MP_COV(do_monster);
    Sprite &sp = e.spr()[si];
    MP_TAB int16_t Gx[5] = { 2, 2, -2, -2, 0 }, Gy[5] = { -2, 2, 2, -2, 0 };
    MP_TAB int16_t ND1[4] = { 0, 1, 2, 3 }, ND2[4] = { 1, 2, 3, 0 };
    MP_TAB int16_t nn1[4] = { 2, 5, 8, 11 }, nn2[4] = { 11, 2, 5, 8 };
    int16_t d, z, c;
    if (sp.soundCount > 0) sp.soundCount--;
    if (sp.control < 0) {
        if (sp.control == -2) {
            d = (int16_t)((sp.frame - 1) / 3);
            z = (int16_t)((sp.frame - 1) % 3);
            if (z == 2) sp.step = 0;
            if (z == 0) sp.step = 1;
            if (sp.step) z++; else z--;
            c = get_dir(e, sp.x, sp.y, sp.destX, sp.destY);
            if (S.absDist < 18) {
                sp.control = -1; sp.count = 1000; sp.flag = 1;
                sp.destX = sp.origX; sp.destY = sp.origY;
            } else {
                c = (int16_t)((c - 1) / 2);
                if ((c != d && rnd(e, 5) == 0) || rnd(e, 20) == 0) {
                    int diff = (c - d) & 3;
                    if (diff == 1 || diff == 3) d = c;
                    else {
                        if (rnd16(e) & 1) d++; else d--;
                        d &= 3;
                    }
                } else if (rnd(e, 20) == 0) {
                    if (rnd16(e) & 1) d++; else d--;
                    d &= 3;
                }
            }
        } else {
            d = (int16_t)((sp.frame - 1) / 3);
            if (d < 4) {
                z = (int16_t)((sp.frame - 1) % 3);
                if (z == 2) sp.step = 0;
                if (z == 0) sp.step = 1;
                if (sp.step) z++; else z--;
                get_dir(e, sp.x, sp.y, sp.destX, sp.destY);
                if (S.absDist < 60) {
                    if (sp.flag == 0) {
                        sp.flag = 1; sp.destX = sp.origX; sp.destY = sp.origY;
                    } else {
                        sp.frame = 0;
                        return;
                    }
                }
                c = get_dir(e, sp.x, sp.y, sp.destX, sp.destY);
                c = (int16_t)((c - 1) / 2);
                if ((c != d) && (!rnd(e, 10))) {
                    if (rnd16(e) & 1) z = ND1[d]; else z = ND2[d];
                    d = 4;
                    if (!sp.soundCount) sp.soundCount = (int16_t)(50 + rnd(e, 100));
                }
            } else {
                d = 4;
                c = sp.frame;
                z = (int16_t)((c - 13) & 3);
                if (!(rnd16(e) & 3)) {
                    if (rnd16(e) & 1) z = nn1[z]; else z = nn2[z];
                    d = (int16_t)((z - 1) / 3);
                    z = (int16_t)((z - 1) % 3);
                }
            }
        }
    } else {
        d = sp.control;
        z = (int16_t)((sp.frame - 1) % 3);
        if (z == 2) sp.step = 0;
        if (z == 0) sp.step = 1;
        if (sp.step) z++; else z--;
    }
    z = (int16_t)(d * 3 + z + 1);
    if (z > 16) z = 16;
    sp.frame = z;
    sp.x = (int16_t)(sp.x + Gx[d]);
    sp.y = (int16_t)(sp.y + Gy[d]);
    if (sp.count > 0) sp.count--;
    c = get_char(e, sp.x + sp.xHot, sp.y + sp.yHot);
    if (c == -1 || (c == T_RIVER && sp.count != 0 && sp.control == -1)) sp.frame = 0;
    hit_vehicles(e, si);
    destroy_map_tile(e, e.spr()[si].x + 48, e.spr()[si].y + 16);
}

// [LEAD] sprite.cpp:1188-1248 doTornadoSprite
MPS void do_tornado(Env &e, int si)
{
    MP_COV(do_tornado);
    Sprite &sp = e.spr()[si];
    MP_TAB int16_t CDx[9] = { 2, 3, 2, 0, -2, -3, 0, 0, 0 }, CDy[9] = { -2, 0, 2, 3, 2, 0, 0, 0, 0 };
    int16_t z = sp.frame;
    if (z == 2) z = sp.flag ? 3 : 1;
    else {
        sp.flag = (z == 1) ? 1 : 0;
        z = 2;
    }
    if (sp.count > 0) sp.count--;
    sp.frame = z;
    hit_vehicles(e, si);
    z = rnd(e, 5);
    Sprite &sq = e.spr()[si];
    sq.x = (int16_t)(sq.x + CDx[z]);
    sq.y = (int16_t)(sq.y + CDy[z]);
    if (sprite_not_in_bounds(sq)) sq.frame = 0;
    if (sq.count != 0 && rnd(e, 500) == 0) sq.frame = 0;
    destroy_map_tile(e, sq.x + 48, sq.y + 40);
}

// [LEAD] sprite.cpp:1255-1288 doExplosionSprite
MPN void do_explosion(Env &e, int si)
{
    MP_COV(do_explosion);
    Sprite &sp = e.spr()[si];
    if ((S.spriteCycle & 1) == 0) sp.frame++;
    if (sp.frame > 6) {
        sp.frame = 0;
        start_fire(e, sp.x + 48 - 8, sp.y + 16);
        start_fire(e, sp.x + 48 - 24, sp.y);
        start_fire(e, sp.x + 48 + 8, sp.y);
        start_fire(e, sp.x + 48 - 24, sp.y + 32);
        start_fire(e, sp.x + 48 + 8, sp.y + 32);
    }
}

// [LEAD] sprite.cpp:548-600 moveObjects.  Every sprite is created with name "" so a sprite whose
// frame is 0 is destroyed when the walk reaches it.  The bus sprite is never generated
// (generateBus call commented out, simulate.cpp:1062).
MPR void move_objects(Env &e)
{
    MP_COV(move_objects);
    if (!S.simSpeed) return;
    S.spriteCycle++;
    MP_NOUNROLL for (int si = S.spriteHead; si >= 0;) {
        if (e.spr()[si].frame > 0) {
            switch (e.spr()[si].type) {
            case SPR_TRAIN: do_train(e, si); break;
            case SPR_HELICOPTER: do_copter(e, si); break;
            case SPR_AIRPLANE: do_airplane(e, si); break;
            case SPR_SHIP: do_ship(e, si); break;
            case SPR_MONSTER: do_monster(e, si); break;
            case SPR_TORNADO: do_tornado(e, si); break;
            case SPR_EXPLOSION: do_explosion(e, si); break;
            default: break;
            }
            si = e.spr()[si].next;
        } else {
            int s = si;
            si = e.spr()[si].next;
            destroy_sprite(e, s);
        }
    }
}

// ------------------------------------------------------------------ sprite generators
MPN void generate_train_make(Env &e, int x, int y)             // sprite.cpp:1831-1837 (guard: generate_train, mp_sim.h)
{
    make_sprite(e, SPR_TRAIN, (x << 4) - 39, (y << 4) + 6);
}

MPN void generate_ship(Env &e)                                 // sprite.cpp:1854-1898
{
    MP_COV(generate_ship);
    if (!(rnd16(e) & 3))
        MP_NOUNROLL for (int x = 4; x < WORLD_W - 2; x++)
            if (e.map()[tix(x, 0)] == T_CHANNEL) { make_sprite(e, SPR_SHIP, (x << 4) - 47, 0); return; }
    if (!(rnd16(e) & 3))
        MP_NOUNROLL for (int y = 1; y < WORLD_H - 2; y++)
            if (e.map()[tix(0, y)] == T_CHANNEL) { make_sprite(e, SPR_SHIP, -47, y << 4); return; }
    if (!(rnd16(e) & 3))
        MP_NOUNROLL for (int x = 4; x < WORLD_W - 2; x++)
            if (e.map()[tix(x, WORLD_H - 1)] == T_CHANNEL) {
                make_sprite(e, SPR_SHIP, (x << 4) - 47, (WORLD_H - 1) << 4);
                return;
            }
    if (!(rnd16(e) & 3))
        MP_NOUNROLL for (int y = 1; y < WORLD_H - 2; y++)
            if (e.map()[tix(WORLD_W - 1, y)] == T_CHANNEL) {
                make_sprite(e, SPR_SHIP, ((WORLD_W - 1) << 4) - 47, y << 4);
                return;
            }
}

MPN void generate_copter(Env &e, int x, int y)                 // sprite.cpp:1968-1975
{
    MP_COV(generate_copter);
    if (get_sprite(e, SPR_HELICOPTER) >= 0) return;
    make_sprite(e, SPR_HELICOPTER, x << 4, (y << 4) + 30);
}

MPN void generate_plane(Env &e, int x, int y)                  // sprite.cpp:1982-1989
{
    MP_COV(generate_plane);
    if (get_sprite(e, SPR_AIRPLANE) >= 0) return;
    make_sprite(e, SPR_AIRPLANE, (x << 4) + 48, (y << 4) + 12);
}

// [LEAD] sprite.cpp:1912-1953 makeMonster / makeMonsterAt
MPN void make_monster(Env &e)
{
    MP_COV(make_monster);
    int si = get_sprite(e, SPR_MONSTER);
    if (si >= 0) {
        e.spr()[si].soundCount = 1; e.spr()[si].count = 1000;
        e.spr()[si].destX = (int16_t)(S.pollutionMaxX << 4);
        e.spr()[si].destY = (int16_t)(S.pollutionMaxY << 4);
        return;
    }
    MP_NOUNROLL for (int z = 0; z < 300; z++) {
        int x = rnd(e, WORLD_W - 20) + 10, y = rnd(e, WORLD_H - 10) + 5;
        int v = e.map()[tix(x, y)];
        if (v == T_RIVER || v == T_RIVER + BULLBIT) {
            make_sprite(e, SPR_MONSTER, (x << 4) + 48, y << 4);
            return;
        }
    }
    make_sprite(e, SPR_MONSTER, (60 << 4) + 48, 50 << 4);
}

// [LEAD] sprite.cpp:1995-2013 makeTornado
MPN void make_tornado(Env &e)
{
    MP_COV(make_tornado);
    int si = get_sprite(e, SPR_TORNADO);
    if (si >= 0) { e.spr()[si].count = 200; return; }
    int16_t x = (int16_t)(rnd(e, (WORLD_W << 4) - 800) + 400);
    int16_t y = (int16_t)(rnd(e, (WORLD_H << 4) - 200) + 100);
    make_sprite(e, SPR_TORNADO, x, y);
}

// ------------------------------------------------------------------ disasters (disasters.cpp)
MPD bool vulnerable(int tem)                                   // disasters.cpp:311-320
{
    MP_COV(vulnerable);
    int t2 = tem & LOMASK;
    return !(t2 < T_RESBASE || t2 > T_LASTZONE || (tem & ZONEBIT));
}

// [LEAD] disasters.cpp:246-269 makeEarthquake
MPN void make_earthquake(Env &e)
{
    MP_COV(make_earthquake);
    int strength = rnd(e, 700) + 300;
    MP_NOUNROLL for (int16_t z = 0; z < strength; z++) {
        int x = rnd(e, WORLD_W - 1), y = rnd(e, WORLD_H - 1);
        if (vulnerable(e.map()[tix(x, y)])) {
            if ((z & 0x3) != 0) mset(e, tix(x, y), (uint16_t)random_rubble(e));
            else mset(e, tix(x, y), (uint16_t)random_fire(e));
        }
    }
}

// [ALL] makeEarthquake with the probes spread over the lanes.  An iteration draws x = getRandom(119) and
// y = getRandom(99) -- two LCG steps unless a draw is rejected (16 resp. 36 of 65536 values) -- and does
// nothing unless the tile is vulnerable (a non-centre building tile: ~0.5 % of a gym city's world).  So
// 32 iterations are evaluated at once, lane k on draws 2k+1 and 2k+2 after the current state (jump-ahead);
// the iterations in front of the first lane that sees a rejected draw or a vulnerable tile are no-ops
// and are committed as 2 draws each, that one iteration is replayed by the serial lane, and so on.
MPN void make_earthquake_coop(Env &e)
{
    MP_COV(make_earthquake_coop);
#if MP_WARP_COOP
    int strength = 0;
    if (e.c.lead()) strength = rnd(e, 700) + 300;
    strength = coop_bcast(e, strength);
    e.c.sync();
    const int lane = e.c.tid;
    int z = 0;
    while (z < strength) {
        const uint32_t rng = S.rng;
        const int nb = tmin(32, strength - z);
        uint32_t x0 = lcg_jump(rng, 2 * lane + 1);
        const int r1 = lcg_r16(x0);
        x0 = x0 * 1103515245u + 12345u;
        const int r2 = lcg_r16(x0);
        bool special = false;
        if (lane < nb) {
            special = r1 >= 65520 || r2 >= 65500;           // getRandom's rejection: (0xffff / range) * range
            if (!special) special = vulnerable(e.map()[tix(r1 % WORLD_W, r2 % WORLD_H)]);
        }
        const unsigned sm = __ballot_sync(0xffffffffu, special);
        const int first = sm ? __ffs(sm) - 1 : nb;
        e.c.sync();
        if (e.c.lead()) {
            S.rng = lcg_jump(rng, 2 * first);
            if (first < nb) {                               // the literal iteration z + first
                const int x = rnd(e, WORLD_W - 1), y = rnd(e, WORLD_H - 1);
                if (vulnerable(e.map()[tix(x, y)])) {
                    if (((z + first) & 0x3) != 0) mset(e, tix(x, y), (uint16_t)random_rubble(e));
                    else mset(e, tix(x, y), (uint16_t)random_fire(e));
                }
            }
        }
        z += first + (first < nb ? 1 : 0);
        e.c.sync();
    }
#else
    make_earthquake(e);
#endif
}

// [LEAD] disasters.cpp:273-290 setFire
MPN void set_fire(Env &e)
{
    MP_COV(set_fire);
    int x = rnd(e, WORLD_W - 1), y = rnd(e, WORLD_H - 1);
    int16_t z = (int16_t)e.map()[tix(x, y)];
    if ((z & ZONEBIT) == 0) {
        z = z & LOMASK;
        if (z > T_LHTHR && z < T_LASTZONE) mset(e, tix(x, y), (uint16_t)random_fire(e));
    }
}

// [LEAD] disasters.cpp:327-368 makeFlood
MPN void make_flood(Env &e)
{
    MP_COV(make_flood);
    MP_TAB int8_t dx[4] = { 0, 1, 0, -1 }, dy[4] = { -1, 0, 1, 0 };
    MP_NOUNROLL for (int z = 0; z < 300; z++) {
        int x = rnd(e, WORLD_W - 1), y = rnd(e, WORLD_H - 1);
        int c = e.map()[tix(x, y)] & LOMASK;
        if (c > T_CHANNEL && c <= T_WATER_HIGH) {
            MP_NOUNROLL for (int t = 0; t < 4; t++) {
                int xx = x + dx[t], yy = y + dy[t];
                if (inb(xx, yy)) {
                    int16_t cc = (int16_t)e.map()[tix(xx, yy)];
                    if (cc == T_DIRT || (cc & (BULLBIT | BURNBIT)) == (BULLBIT | BURNBIT)) {
                        mset(e, tix(xx, yy), T_FLOOD);
                        S.floodCount = 30;
                        return;
                    }
                }
            }
        }
    }
}

// [LEAD] disasters.cpp:82-144 doDisasters (scenario == SC_NONE on this path, so
// disasterEvent never leaves SC_NONE and scenarioDisaster is unreachable)
// defer_quake: return 1 instead of running makeEarthquake (the caller runs make_earthquake_coop on all lanes)
MPN int do_disasters(Env &e, bool defer_quake)
{
    MP_COV(do_disasters);
    MP_TAB int16_t chance[3] = { 10 * 48, 5 * 48, 60 };
    if (S.floodCount) S.floodCount--;
    if (!S.enableDisasters) return 0;
    int lvl = S.gameLevel;
    if (lvl > 2) lvl = 0;
    if (!rnd(e, chance[lvl])) {
        switch (rnd(e, 8)) {
        case 0: case 1: set_fire(e); break;
        case 2: case 3: make_flood(e); break;
        case 4: break;
        case 5: make_tornado(e); break;
        case 6:
            if (defer_quake) return 1;
            make_earthquake(e);
            break;
        case 7: case 8:
            if (S.pollutionAverage > 60) make_monster(e);
            break;
        }
    }
    return 0;
}

#undef S
} // namespace mp
