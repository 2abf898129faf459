// obake/b200/engine.hpp -- C++ RAII face of the C ABI (include/obk_b200.h) + coefficient marshalling traits.
//
// Status codes are mapped onto the exception types the reference throws on this path
// (polynomial.hpp:847-850, 1544-1550): std::overflow_error for exponent / table / coefficient overflow,
// std::invalid_argument for malformed operands, std::runtime_error for CUDA failures.  There is no CPU
// fallback: without the library or a B200 the constructor throws.
#ifndef OBAKE_B200_ENGINE_HPP
#define OBAKE_B200_ENGINE_HPP

#include <cstdint>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../../../include/obk_b200.h"
#include "../integer.hpp"

namespace obake::b200
{

class engine
{
    obk_engine *m_h = nullptr;
    std::mutex m_mtx; // calls are serialised per handle (SURVEY 8b "Threading")
    unsigned m_ndev = 1;

public:
    // Single-process multi-device engine (obk_engine_create_multi): obake::operator* and truncated_mul spread over
    // the devices; device-resident chains (device_polynomial) need a one-device engine.
    explicit engine(const std::vector<int> &devices) : m_ndev(static_cast<unsigned>(devices.size()))
    {
        const auto st = ::obk_engine_create_multi(devices.data(), static_cast<int>(devices.size()), &m_h);
        if (st != OBK_OK) {
            throw std::runtime_error(std::string("obk_engine_create_multi failed: ") + ::obk_last_error(nullptr)
                                     + " (the B200 engine has no CPU fallback)");
        }
    }
    unsigned n_devices() const
    {
        return m_ndev;
    }
    explicit engine(int device = 0)
    {
        const auto st = ::obk_engine_create(device, &m_h);
        if (st != OBK_OK) {
            throw std::runtime_error(std::string("obk_engine_create failed: ") + ::obk_last_error(nullptr)
                                     + " (the B200 engine has no CPU fallback)");
        }
    }
    engine(const engine &) = delete;
    engine &operator=(const engine &) = delete;
    ~engine()
    {
        ::obk_engine_destroy(m_h);
    }
    obk_engine *handle()
    {
        return m_h;
    }
    std::mutex &mutex()
    {
        return m_mtx;
    }
    void set_option(const char *name, std::int64_t v)
    {
        check(::obk_engine_set_option(m_h, name, v));
    }
    void check(obk_status st) const
    {
        if (st == OBK_OK) {
            return;
        }
        const std::string msg = ::obk_last_error(m_h);
        switch (st) {
            case OBK_ERR_KEY_OVERFLOW:
            case OBK_ERR_TABLE_OVERFLOW:
            case OBK_ERR_CF_OVERFLOW:
                throw std::overflow_error(msg);
            case OBK_ERR_INVALID_ARG:
            case OBK_ERR_UNSUPPORTED:
                throw std::invalid_argument(msg);
            case OBK_ERR_NOMEM:
                throw std::bad_alloc();
            default:
                throw std::runtime_error(msg);
        }
    }
    // Process-wide default engine, created on first use: the devices listed in OBAKE_B200_DEVICES ("0,1,2,3": the
    // single-process multi-device engine), else device OBAKE_B200_DEVICE (default 0).
    static engine &instance()
    {
        static std::unique_ptr<engine> inst;
        static std::once_flag flag;
        std::call_once(flag, [] {
            if (const char *l = std::getenv("OBAKE_B200_DEVICES")) {
                std::vector<int> devs;
                for (const char *q = l; *q;) {
                    char *end = nullptr;
                    const long v = std::strtol(q, &end, 10);
                    if (end == q) {
                        break;
                    }
                    devs.push_back(static_cast<int>(v));
                    q = *end == ',' ? end + 1 : end;
                }
                if (devs.size() > 1u) {
                    inst = std::make_unique<engine>(devs);
                    return;
                }
                if (devs.size() == 1u) {
                    inst = std::make_unique<engine>(devs[0]);
                    return;
                }
            }
            int dev = 0;
            if (const char *s = std::getenv("OBAKE_B200_DEVICE")) {
                dev = std::atoi(s);
            }
            inst = std::make_unique<engine>(dev);
        });
        return *inst;
    }
};

// How a coefficient type crosses the ABI.
template <typename C, typename = void>
struct cf_traits;

template <>
struct cf_traits<double> {
    static constexpr std::uint32_t kind = OBK_CF_F64;
    static unsigned limbs_needed(const double &)
    {
        return 1;
    }
    static void store(const double &c, std::uint64_t *out, unsigned)
    {
        static_assert(sizeof(double) == 8);
        __builtin_memcpy(out, &c, 8);
    }
    static double load(const std::uint64_t *in, unsigned)
    {
        double d;
        __builtin_memcpy(&d, in, 8);
        return d;
    }
};

template <>
struct cf_traits<::obake::integer> {
    static constexpr std::uint32_t kind = OBK_CF_INT;
    static unsigned limbs_needed(const ::obake::integer &c)
    {
        return c.limbs_needed();
    }
    static void store(const ::obake::integer &c, std::uint64_t *out, unsigned n)
    {
        c.to_limbs(out, n);
    }
    static ::obake::integer load(const std::uint64_t *in, unsigned n)
    {
        return ::obake::integer::from_limbs(in, n);
    }
};

} // namespace obake::b200

#endif
