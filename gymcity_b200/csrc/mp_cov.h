// mp_cov.h -- function-hit coverage of the device build (-DMP_COVERAGE; tests/test_coverage_gpu.py).
// Every engine / gym function starts with MP_COV(name): a no-op in the product build and on the host; in the coverage
// build it sets bit COV_name of a device-global bitmap that mpb_coverage() reads back.  The list is generated
// (tools/gen_cov_list.py) from the function definitions in mp_sim.h, mp_sprites.h, mp_tools.h and mp_gym.h.
#pragma once
#define MP_COV_LIST(X) \
    X(set_valves) X(clear_census_scalars) X(push_power_stack) X(set_zone_power) X(road_test) X(find_perimeter_road) \
    X(tile_toward) X(try_go) X(drive_done) X(try_drive) X(add_to_traffic_density) X(make_traffic) \
    X(inc_rate_of_growth) X(land_pollution_value) X(zone_plop) X(do_free_pop) X(eval_lot) X(build_house) \
    X(res_plop) X(com_plop) X(ind_plop) X(do_res_in) X(do_res_out) X(make_hospital) \
    X(eval_res) X(do_residential) X(do_commercial) X(set_smoke) X(do_ind_out) X(do_industrial) \
    X(repair_zone) X(do_hospital_church) X(do_meltdown) X(do_special_zone) X(do_zone) X(fire_zone) \
    X(do_fire) X(do_flood) X(do_bridge) X(do_road) X(do_rail) X(scan_tile) \
    X(tile_kind) X(lcg_jump) X(scan_word_batch) X(map_scan_dense) X(map_scan) X(scroll_history) \
    X(take10_census) X(take120_census) X(do_budget) X(collect_tax) X(dec_traffic_cell) X(dec_traffic_map) \
    X(dec_rog_map) X(city_class_of) X(send_messages) X(power_scan_walk) X(power_scan_parallel) X(do_power_scan) \
    X(zero_byte2) X(smooth_byte2_box) X(box_add) X(box_reduce) X(smooth_station) X(city_center_distance) \
    X(pollution_value) X(coop_add) X(tie_max_scan) X(ptlv_scan) X(crime_scan) X(population_density_of) \
    X(pop_density_scan) X(com_rate_map) X(fire_analysis) X(city_evaluation) X(coop_bcast) X(scan_due) \
    X(simulate_phase) X(simulate) X(sim_phase) X(sim_loop_head) X(sim_loop_tail) X(sim_loop) \
    X(do_sim_init) X(animate_tiles) X(get_sprite) X(init_sprite) X(new_sprite) X(destroy_sprite) \
    X(make_sprite) X(make_explosion) X(boat_distance) X(get_char) X(turn_to) X(try_other) \
    X(sprite_not_in_bounds) X(get_dir) X(sprite_collision) X(explode_sprite) X(start_fire_in_zone) X(destroy_map_tile) \
    X(start_fire) X(do_train) X(do_copter) X(do_airplane) X(do_ship) X(hit_vehicles) \
    X(do_monster) X(do_tornado) X(do_explosion) X(move_objects) X(generate_train) X(generate_ship) \
    X(generate_copter) X(generate_plane) X(make_monster) X(make_tornado) X(vulnerable) X(make_earthquake) \
    X(make_earthquake_coop) X(set_fire) X(make_flood) X(do_disasters) X(tally) X(fix_single) \
    X(fix_candidate) X(fix_zone) X(lay_doze) X(lay_road) X(lay_rail) X(lay_wire) \
    X(connect_tile) X(check_big_zone) X(check_size) X(put_rubble) X(bulldozer_tool) X(build_building) \
    X(smooth_trees_at_fx) X(tool_down) X(put_on_map) X(plop_b_river) X(plop_s_river) X(do_river) \
    X(smooth_river) X(smooth_trees) X(do_trees) X(make_naked_island) X(generate_map) X(gen_put) \
    X(gen_plop_b) X(gen_plop_s) X(gen_river) X(gen_redge_value) X(gen_smooth_river) X(gen_smooth_trees) \
    X(gen_trees) X(generate_map_coop) X(gym_zone_of_tile) X(gym_update_tile) X(gym_clear_tile) X(gym_clear_patch) \
    X(gym_add_zone) X(gym_add_zone_bot) X(gym_apply_action) X(gym_paint_action) X(gym_init_age_array) X(gym_extinguish) \
    X(gym_set_empty) X(gym_update_map) X(gym_city_metrics) X(gym_reward) X(gym_observe) X(gym_step_pre) \
    X(gym_step_post) X(gym_step) X(gym_reset_begin) X(gym_reset_post) X(gym_reset_finish)

namespace mp {
#define MP_COV_ENUM(n) COV_##n,
enum CovId : int { MP_COV_LIST(MP_COV_ENUM) COV_COUNT };
#undef MP_COV_ENUM
constexpr int COV_WORDS = (COV_COUNT + 63) / 64;
}
#if defined(MP_COVERAGE) && defined(__CUDA_ARCH__)
extern __device__ unsigned long long mp_cov_bits[mp::COV_WORDS];
#define MP_COV(n) do { const int id_ = mp::COV_##n; if (!((mp_cov_bits[id_ >> 6] >> (id_ & 63)) & 1ull)) atomicOr(&mp_cov_bits[id_ >> 6], 1ull << (id_ & 63)); } while (0)
#else
#define MP_COV(n) ((void)0)
#endif
