/*! \file common.hpp
 *  \brief Declarations shared by the `cl_algo::ICP` classes (counterpart of /root/reference/include/ICP/common.hpp:43-49).
 */
#ifndef ICP_COMMON_HPP
#define ICP_COMMON_HPP

#include <cstdint>

namespace cl_algo
{
namespace ICP
{
    /*! \brief Which (pinned) staging buffers `init` instantiates. */
    enum class Staging : uint8_t
    {
        NONE,  /*!< no staging buffers: write / read are no-ops returning nullptr */
        I,     /*!< input staging buffers */
        O,     /*!< output staging buffers */
        IO     /*!< both */
    };
}
}

#endif  // ICP_COMMON_HPP
